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
  int groups = 1;                  // block_ws_kernel: depthwise warp groups (tiles in flight on the CUDA cores)
  int dw_threads = 0;              // block_ws_kernel: threads per depthwise group
  int out_bufs = 1;                // block_ws_kernel: output staging buffers (TMA store of tile i overlaps the epilogue of i+1); 0: none, the
                                   // epilogue stores to `out_direct`
  float* out_direct = nullptr;
  int acc_cols = 32;               // block_ws_kernel: TMEM columns per accumulator buffer
  int in_pad = 0;                  // block_ws_kernel: input tile pixel stride padded to an odd number of quads
  int skip_tma = 0;                // block_ws_kernel, two epilogue teams: the residual tile (skip_mode 2) is loaded by TMA into the staging buffer
  int tc_cp = 0, tc_np = 0;        // blaze_block_tc_kernel: pixel strides (floats) of the input / output staging tiles (>= C / N)
  // block_ws_kernel, f16-split mode: the depthwise result is written as ONE plane set of (f16 hi, f16 lo) pairs and multiplied
  // with tcgen05 kind::f16 (half the A-operand bytes in shared memory of the tf32 hi / lo planes, same 2^-22 fidelity)
  int f16 = 0;
  int wsplit16 = 1;                // 1: weights f16-exact; 2: + (w - f16(w)) against the hi half
  const float* w_f16 = nullptr;    // [wsplit16][C/4][Np][8 halves] (Step::w_f16)
  float bias_c[128] = {};          // f16 mode: pointwise bias by value (added in the epilogue, not as a K step)
  const int* n_active = nullptr;
  float alpha_c[128] = {};         // block_ws_kernel: PRELU slopes by value (read through the constant bank in the epilogue)
};

struct BlockTcLaunch {
  BlockTcArgs args;
  const float* alpha_host = nullptr;   // host copy of the PRELU slopes [N] (block_ws_kernel passes them as kernel parameters)
  const float* bias_host = nullptr;    // host copy of the pointwise bias [N] (block_ws_kernel, f16 mode)
  const float* in = nullptr;       // [B,H,W,C]
  float* out = nullptr;            // [B,H,W,N]
};

// ---- general tensor-core convolution (conv_tc_kernel.cu) ----
struct TViewC { const float* p = nullptr; long long bstride = 0; int H = 0, W = 0, C = 0; };
struct TViewM { float* p = nullptr; long long bstride = 0; int H = 0, W = 0, C = 0; };
struct ConvTcArgs {
  TViewC in, skip;
  TViewM out;
  int mode = 0;                    // 0: CONV_2D (im2col gather), 1: BlazeBlock (depthwise 3x3 evaluated in the gather)
  int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;
  int K = 0, Kp = 0, N = 0, Nt = 0, n_tiles = 1;
  const float* w_tc = nullptr;     // [n_tiles][wsplit][Kp/4][Nt][4]
  const float* bias = nullptr;
  const float* w_dw = nullptr;
  const float* b_dw = nullptr;
  const float* alpha = nullptr;
  int act = 0, has_skip = 0, skip_pool = 0, skip_c = 0;
  int wsplit = 1, tmem_cols = 32;
  int B = 0;
  const int* n_active = nullptr;
};
cudaError_t conv_tc_init();
bool conv_tc_supported(const Step& s);
cudaError_t launch_conv_tc(const ConvTcArgs& a, cudaStream_t stream);

cudaError_t mma_kernels_init();                             // once per device
bool block_tc_supported(const Step& s);    // can this planned step run on the tensor-core kernel?
cudaError_t launch_block_tc(const BlockTcLaunch& l, cudaStream_t stream);

// ---- warp-specialised stride-1 BlazeBlock kernel (block_ws_kernel.cu) ----
cudaError_t block_ws_init();
bool block_ws_supported(const Step& s);
cudaError_t launch_block_ws(const BlockTcLaunch& l, cudaStream_t stream);

}  // namespace fdl
