// stem_kernel.cu -- the RGB stem convolutions: CONV_2D 5x5 / 3x3, stride 2, Cin = 3, SAME padding, + RELU/PRELU
// (first op of every graph; 59 MFLOP of the back-256 detector's 378).
//
// K = kh*kw*3 is tiny (27 / 75) and the arithmetic intensity is ~25 flop/B, so this one is FFMA-bound rather
// than HBM-bound: each CTA stages the input patch of an 8 x TW output tile and the whole [K][Cout] weight matrix
// in shared memory; each thread keeps 4 horizontally adjacent output pixels x 8 output channels in registers
// (32 accumulators), loads the 4 pixels' shared input row segment once per kernel row as float4s, and reads the
// weights as warp-uniform (broadcast) float4s.
#include <cuda_runtime.h>

#include "net_kernels.cuh"
#include "plan.h"

namespace fdl {

namespace {

constexpr int kTileH = 8;

template <int KH, int KW>
__global__ void __launch_bounds__(512) stem_conv_kernel(const ConvArgs a, const int TWo, const int tiles_x, const int tiles_y) {
  extern __shared__ __align__(16) float sm[];
  constexpr int K = KH * KW * 3;
  constexpr int NIN = (6 + KW) * 3;             // input floats one thread needs per kernel row (4 pixels, stride 2)
  constexpr int NV = (NIN + 3) / 4;             // as float4s
  const int N = a.N;                            // multiple of 8
  const int PR = (kTileH - 1) * 2 + KH;         // patch rows
  const int PC = ((TWo - 1) * 2 + KW) * 3;      // patch floats per row
  const int PCp = ((PC + 4 + 3) / 4) * 4;       // padded (threads over-read up to 3 floats)
  float* s_w = sm;                              // [K][N]
  float* s_patch = sm + K * N;                  // [PR][PCp]

  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int tiles_per_img = tiles_x * tiles_y;
  const int b = blockIdx.x / tiles_per_img;
  if (b >= nb) return;
  const int r = blockIdx.x - b * tiles_per_img;
  const int ty = r / tiles_x, tx = r - ty * tiles_x;
  const int tid = threadIdx.x, nthreads = blockDim.x;

  for (int i = tid; i < K * N; i += nthreads) {
    int k = i / N, n = i - k * N;
    s_w[i] = __ldg(a.w + (long long)k * a.Npad + n);
  }
  {
    const int iy0 = ty * kTileH * 2 - a.pad_t;
    const int ic0 = (tx * TWo * 2 - a.pad_l) * 3;     // first float of the patch row inside the image row
    const int row_floats = a.in.W * 3;
    const float* src = a.in.p + (long long)b * a.in.bstride;
    for (int i = tid; i < PR * PCp; i += nthreads) {
      int pr = i / PCp, pc = i - pr * PCp;
      int iy = iy0 + pr, ic = ic0 + pc;
      float v = 0.f;
      if (pc < PC && iy >= 0 && iy < a.in.H && ic >= 0 && ic < row_floats) v = __ldg(src + (long long)iy * row_floats + ic);
      s_patch[i] = v;
    }
  }
  __syncthreads();

  const int G = kTileH * (TWo / 4);             // pixel groups per CTA (multiple of 32: the channel group is warp-uniform)
  const int g = tid % G, cg = tid / G;
  const int gy = g / (TWo / 4), gx = g - gy * (TWo / 4);
  const int n0 = cg * 8;

  float acc[4][8];
  {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + n0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 4));
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      acc[p][0] = b0.x; acc[p][1] = b0.y; acc[p][2] = b0.z; acc[p][3] = b0.w;
      acc[p][4] = b1.x; acc[p][5] = b1.y; acc[p][6] = b1.z; acc[p][7] = b1.w;
    }
  }
#pragma unroll
  for (int ky = 0; ky < KH; ++ky) {
    const float* rowp = s_patch + (gy * 2 + ky) * PCp + gx * 24;     // 4 pixels * stride 2 * 3 channels = 24 floats per group
    float in[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 t = *reinterpret_cast<const float4*>(rowp + 4 * v);
      in[4 * v] = t.x; in[4 * v + 1] = t.y; in[4 * v + 2] = t.z; in[4 * v + 3] = t.w;
    }
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* wp = s_w + ((ky * KW + kx) * 3 + c) * N + n0;
        float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float x = in[(2 * p + kx) * 3 + c];
          acc[p][0] = fmaf(x, w0.x, acc[p][0]); acc[p][1] = fmaf(x, w0.y, acc[p][1]);
          acc[p][2] = fmaf(x, w0.z, acc[p][2]); acc[p][3] = fmaf(x, w0.w, acc[p][3]);
          acc[p][4] = fmaf(x, w1.x, acc[p][4]); acc[p][5] = fmaf(x, w1.y, acc[p][5]);
          acc[p][6] = fmaf(x, w1.z, acc[p][6]); acc[p][7] = fmaf(x, w1.w, acc[p][7]);
        }
      }
    }
  }
  const int oy = ty * kTileH + gy;
  if (oy >= a.out.H) return;
  float al[8];
  if (a.act == ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) al[j] = __ldg(a.alpha + n0 + j);
  }
  float* orow = a.out.p + (long long)b * a.out.bstride + (long long)oy * a.out.W * N;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int ox = tx * TWo + gx * 4 + p;
    if (ox >= a.out.W) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = acc[p][j];
      if (a.act == ACT_RELU) t = fmaxf(t, 0.f);
      else if (a.act == ACT_PRELU) t = t >= 0.f ? t : t * al[j];
      v[j] = t;
    }
    float4* op = reinterpret_cast<float4*>(orow + (long long)ox * N + n0);
    op[0] = make_float4(v[0], v[1], v[2], v[3]);
    op[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

struct StemCfg { int TWo, threads, tiles_x, tiles_y; size_t smem; };

StemCfg stem_cfg(const ConvArgs& a) {
  StemCfg c;
  const int cgs = a.N / 8;
  // widest tile that keeps the CTA at <= 512 threads; prefer the one that wastes fewer columns
  c.TWo = (a.out.W % 64 == 0 && kTileH * 16 * cgs <= 512) ? 64 : 32;
  c.threads = kTileH * (c.TWo / 4) * cgs;
  c.tiles_x = (a.out.W + c.TWo - 1) / c.TWo;
  c.tiles_y = (a.out.H + kTileH - 1) / kTileH;
  const int K = a.kh * a.kw * 3;
  const int PR = (kTileH - 1) * 2 + a.kh;
  const int PC = ((c.TWo - 1) * 2 + a.kw) * 3;
  const int PCp = ((PC + 4 + 3) / 4) * 4;
  c.smem = (size_t)(K * a.N + PR * PCp) * sizeof(float);
  return c;
}

}  // namespace

bool stem_supported(const ConvArgs& a) {
  if (a.mode != 0 || a.in.C != 3 || a.stride != 2 || a.kh != a.kw || (a.kh != 3 && a.kh != 5)) return false;
  if (a.N % 8 != 0 || a.N > 64 || a.has_skip) return false;
  if (a.out.bstride != (long long)a.out.H * a.out.W * a.N) return false;
  StemCfg c = stem_cfg(a);
  return c.threads <= 512 && c.threads % 32 == 0 && c.smem <= 96 * 1024;
}

cudaError_t stem_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel<5, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(stem_conv_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

cudaError_t launch_stem_conv(const ConvArgs& a, cudaStream_t stream) {
  StemCfg c = stem_cfg(a);
  const unsigned grid = (unsigned)(a.B * c.tiles_x * c.tiles_y);
  if (grid == 0) return cudaSuccess;
  if (a.kh == 5) stem_conv_kernel<5, 5><<<grid, c.threads, c.smem, stream>>>(a, c.TWo, c.tiles_x, c.tiles_y);
  else stem_conv_kernel<3, 3><<<grid, c.threads, c.smem, stream>>>(a, c.TWo, c.tiles_x, c.tiles_y);
  count_launch();
  return cudaGetLastError();
}

}  // namespace fdl
