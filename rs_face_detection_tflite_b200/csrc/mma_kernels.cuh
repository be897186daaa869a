// mma_kernels.cuh -- launch interface of the tensor-core BlazeBlock kernel (see mma_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "plan.h"

namespace fdl {

struct BlockTcArgs {
  const float* w_umma = nullptr;   // [wsplit][C/4][Np][4] pointwise weights in UMMA K-major core-matrix order
  const float* bias = nullptr;     // [N]
  const float* w_dw = nullptr;     // [9][C]
  const float* b_dw = nullptr;     // [C]
  const float* alpha = nullptr;    // [N] or null
  const float* skip = nullptr;     // residual source in global memory (skip_mode 2 / 3)
  long long skip_bstride = 0;
  int C = 0, N = 0, Np = 0;        // Cin (== K), Cout, Cout rounded up to 16
  int H = 0, W = 0, B = 0;
  int tiles_x = 0, tiles_y = 0;
  int act = 0;
  int stride = 1;                  // depthwise stride (1: SAME pad 1; 2: SAME pad 0 before / 1 after on even sizes)
  int pad = 1;
  int alias_out = 0;               // output staging tile shares the A planes (large channel counts)
  int skip_mode = 0;               // 0 none, 1 identity from the resident input tile, 2 direct from global, 3 MAX_POOL 2x2 from global,
                                   // 4 MAX_POOL 2x2 from the resident input tile (stride-2 blocks)
  int skip_c = 0;                  // channels of the residual source (< N: zero channel PAD)
  int stages = 2;                  // input tile buffers
  int wsplit = 1;                  // 1: weights are tf32-exact (A split only); 2: W_hi + W_lo
  int tmem_cols = 32;
  const int* n_active = nullptr;
};

struct BlockTcLaunch {
  BlockTcArgs args;
  const float* in = nullptr;       // [B,H,W,C]
  float* out = nullptr;            // [B,H,W,N]
};

cudaError_t mma_kernels_init();                             // once per device
bool block_tc_supported(const Step& s);    // can this planned step run on the tensor-core kernel?
cudaError_t launch_block_tc(const BlockTcLaunch& l, cudaStream_t stream);

}  // namespace fdl
