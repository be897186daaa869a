// stem_kernel.cu -- the RGB stem convolutions: CONV_2D 5x5 / 3x3, stride 2, Cin = 3, SAME padding, + RELU/PRELU
// (first op of every graph; 59 MFLOP of the back-256 detector's 378).
//
// K = kh*kw*3 is tiny (27 / 75) and the arithmetic intensity is ~25 flop/B, so this one is FFMA-bound rather
// than HBM-bound: each CTA stages the input patch of an 8 x TW output tile and the whole [K][Cout] weight matrix
// in shared memory; each thread keeps 4 horizontally adjacent output pixels x 8 output channels in registers
// (32 accumulators), loads the 4 pixels' shared input row segment once per kernel row as float4s, and reads the
// weights as warp-uniform (broadcast) float4s.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "net_kernels.cuh"
#include "pdl.h"
#include "plan.h"

namespace fdl {

namespace {

constexpr int kTileH = 8;
typedef unsigned long long ull;

__device__ __forceinline__ ull fma2(ull a, ull b, ull c) {
  ull d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ ull pack2(float lo, float hi) {
  ull d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float unpack_lo(ull v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float unpack_hi(ull v) { return __uint_as_float((unsigned)(v >> 32)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;               // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs: the [K][Cout] weight matrix is staged once per CTA; the input patch of tile i+1 streams into the
// second patch buffer with 16-byte cp.async (zero-filled outside the image == SAME padding) while tile i is computed.
// `shift` floats are prepended to every patch row so that the 16-byte chunks are aligned in global memory.
template <int KH, int KW, int SH>
__global__ void __launch_bounds__(512) stem_conv_kernel(const ConvArgs a, const int TWo, const int tiles_x, const int tiles_y) {
  constexpr int shift = SH;
  extern __shared__ __align__(16) float sm[];
  constexpr int K = KH * KW * 3;
  constexpr int NIN = (6 + KW) * 3;             // input floats one thread needs per kernel row (4 pixels, stride 2)
  constexpr int NV = (NIN + SH + 3) / 4;        // as float4s (the thread's segment starts SH floats into an aligned quad)
  const int N = a.N;                            // multiple of 8
  const int PR = (kTileH - 1) * 2 + KH;         // patch rows
  const int PC = ((TWo - 1) * 2 + KW) * 3;      // patch floats per row
  const int NCH = (PC + shift + 3) / 4;         // 16-byte chunks per patch row
  const int PCp = NCH * 4 + 4;                  // padded (threads over-read up to 3 floats)
  float* s_w = sm;                              // [K][N]
  float* s_patch0 = sm + ((K * N + 3) & ~3);    // 2 x [PR][PCp]

  const int tid = threadIdx.x, nthreads = blockDim.x;
  // the weights do not depend on the previous launch: staged before the PDL wait (see pdl.h)
  for (int i = tid; i < K * N; i += nthreads) {
    int k = i / N, n = i - k * N;
    s_w[i] = __ldg(a.w + (long long)k * a.Npad + n);
  }
  pdl_launch_dependents();
  pdl_wait();
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int tiles_per_img = tiles_x * tiles_y;
  const int total = nb * tiles_per_img;
  if ((int)blockIdx.x >= total) return;

  const int row_floats = a.in.W * 3;
  auto issue_patch = [&](int tile, int buf) {
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int iy0 = ty * kTileH * 2 - a.pad_t;
    const int ic0 = (tx * TWo * 2 - a.pad_l) * 3 - shift;   // multiple of 4: chunks never straddle the row ends
    const float* src = a.in.p + (long long)b * a.in.bstride;
    float* dst = s_patch0 + buf * PR * PCp;
    for (int i = tid; i < PR * NCH; i += nthreads) {
      const int pr = i / NCH, ch = i - pr * NCH;
      const int iy = iy0 + pr, ic = ic0 + 4 * ch;
      const bool ok = iy >= 0 && iy < a.in.H && ic >= 0 && ic + 3 < row_floats;
      cp_async16(dst + pr * PCp + 4 * ch, ok ? src + (long long)iy * row_floats + ic : a.in.p, ok);
    }
    cp_async_commit();
  };
  issue_patch(blockIdx.x, 0);

  const int G = kTileH * (TWo / 4);             // pixel groups per CTA (multiple of 32: the channel group is warp-uniform)
  const int g = tid % G, cg = tid / G;
  const int gy = g / (TWo / 4), gx = g - gy * (TWo / 4);
  const int n0 = cg * 8;
  const float4 bias0 = __ldg(reinterpret_cast<const float4*>(a.bias + n0)), bias1 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 4));
  float al[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) al[j] = a.act == ACT_PRELU ? __ldg(a.alpha + n0 + j) : 0.f;

  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = tile + gridDim.x;
    if (next < total) { issue_patch(next, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                            // this tile's patch (and, first time, the weights) are visible to every thread
    const float* s_patch = s_patch0 + buf * PR * PCp;
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;

    // 32 accumulators as 16 packed pairs: fma.rn.f32x2 (FFMA2) does two FMAs per issue slot -- the same arithmetic, half the
    // issue pressure (this kernel was issue-bound: 73 % of its instructions were FFMAs at 40 % FMA-pipe utilisation)
    ull acc2[4][4];
    {
      const ull b01 = pack2(bias0.x, bias0.y), b23 = pack2(bias0.z, bias0.w), b45 = pack2(bias1.x, bias1.y), b67 = pack2(bias1.z, bias1.w);
#pragma unroll
      for (int p = 0; p < 4; ++p) { acc2[p][0] = b01; acc2[p][1] = b23; acc2[p][2] = b45; acc2[p][3] = b67; }
    }
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const float* rowp = s_patch + (gy * 2 + ky) * PCp + gx * 24;   // 4 pixels * stride 2 * 3 channels = 24 floats per group
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
          const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp), w1 = *reinterpret_cast<const ulonglong2*>(wp + 4);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float x = in[(2 * p + kx) * 3 + c + SH];
            const ull xx = pack2(x, x);
            acc2[p][0] = fma2(xx, w0.x, acc2[p][0]); acc2[p][1] = fma2(xx, w0.y, acc2[p][1]);
            acc2[p][2] = fma2(xx, w1.x, acc2[p][2]); acc2[p][3] = fma2(xx, w1.y, acc2[p][3]);
          }
        }
      }
    }
    const int oy = ty * kTileH + gy;
    if (oy < a.out.H) {
      float* orow = a.out.p + (long long)b * a.out.bstride + (long long)oy * a.out.W * N;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ox = tx * TWo + gx * 4 + p;
        if (ox >= a.out.W) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = (j & 1) ? unpack_hi(acc2[p][j >> 1]) : unpack_lo(acc2[p][j >> 1]);
          if (a.act == ACT_RELU) t = fmaxf(t, 0.f);
          else if (a.act == ACT_PRELU) t = t >= 0.f ? t : t * al[j];
          v[j] = t;
        }
        float4* op = reinterpret_cast<float4*>(orow + (long long)ox * N + n0);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();                            // everyone is done with this patch buffer before it is refilled
  }
}

struct StemCfg { int TWo, threads, tiles_x, tiles_y, shift; size_t smem; };

StemCfg stem_cfg(const ConvArgs& a) {
  StemCfg c;
  const int cgs = a.N / 8;
  // widest tile that keeps the CTA at <= 512 threads; prefer the one that wastes fewer columns
  static const int tw_env = getenv("FDL_STEM_TW") ? atoi(getenv("FDL_STEM_TW")) : 0;   // A/B timing: force 32- or 64-wide tiles
  c.TWo = (a.out.W % 64 == 0 && kTileH * 16 * cgs <= 512) ? 64 : 32;
  if (tw_env == 32 || (tw_env == 64 && a.out.W % 64 == 0 && kTileH * 16 * cgs <= 512)) c.TWo = tw_env;
  c.threads = kTileH * (c.TWo / 4) * cgs;
  c.tiles_x = (a.out.W + c.TWo - 1) / c.TWo;
  c.tiles_y = (a.out.H + kTileH - 1) / kTileH;
  const int K = a.kh * a.kw * 3;
  const int PR = (kTileH - 1) * 2 + a.kh;
  const int PC = ((c.TWo - 1) * 2 + a.kw) * 3;
  c.shift = ((-a.pad_l * 3) % 4 + 4) % 4;          // floats prepended to a patch row so that its 16-byte chunks are aligned
  const int PCp = ((PC + c.shift + 3) / 4) * 4 + 4;
  c.smem = (size_t)(((K * a.N + 3) & ~3) + 2 * PR * PCp) * sizeof(float);
  return c;
}

}  // namespace

bool stem_supported(const ConvArgs& a) {
  if (a.mode != 0 || a.in.C != 3 || a.stride != 2 || a.kh != a.kw || (a.kh != 3 && a.kh != 5)) return false;
  if (a.N % 8 != 0 || a.N > 64 || a.has_skip) return false;
  if (a.out.bstride != (long long)a.out.H * a.out.W * a.N) return false;
  // 16-byte cp.async of the input rows: row length, batch stride and base must be multiples of 4 floats
  if ((a.in.W * 3) % 4 != 0 || a.in.bstride % 4 != 0 || (reinterpret_cast<uintptr_t>(a.in.p) & 15) != 0) return false;
  StemCfg c = stem_cfg(a);
  // instantiated: <5,5,1> (SAME, pad 1), <3,3,0> (SAME on even sizes, pad 0) and <3,3,1> (explicit pad 1: sparse full-range stem)
  if (!((a.kh == 5 && c.shift == 1) || (a.kh == 3 && (c.shift == 0 || c.shift == 1)))) return false;
  return c.threads <= 512 && c.threads % 32 == 0 && c.smem <= 96 * 1024;
}

cudaError_t stem_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel<5, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_conv_kernel<3, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(stem_conv_kernel<3, 3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

cudaError_t launch_stem_conv(const ConvArgs& a, cudaStream_t stream) {
  StemCfg c = stem_cfg(a);
  const long long total = (long long)a.B * c.tiles_x * c.tiles_y;
  if (total == 0) return cudaSuccess;
  int per_sm = 1;
  if (a.kh == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stem_conv_kernel<5, 5, 1>, c.threads, c.smem);
  else if (c.shift == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stem_conv_kernel<3, 3, 1>, c.threads, c.smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stem_conv_kernel<3, 3, 0>, c.threads, c.smem);
  if (per_sm < 1) per_sm = 1;
  const unsigned grid = (unsigned)(total < (long long)persist_sms() * per_sm ? total : (long long)persist_sms() * per_sm);     // persistent CTAs
  cudaError_t e;
  if (a.kh == 5) e = launch_pdl(stem_conv_kernel<5, 5, 1>, dim3(grid), dim3(c.threads), c.smem, stream, a, c.TWo, c.tiles_x, c.tiles_y);
  else if (c.shift == 1) e = launch_pdl(stem_conv_kernel<3, 3, 1>, dim3(grid), dim3(c.threads), c.smem, stream, a, c.TWo, c.tiles_x, c.tiles_y);
  else e = launch_pdl(stem_conv_kernel<3, 3, 0>, dim3(grid), dim3(c.threads), c.smem, stream, a, c.TWo, c.tiles_x, c.tiles_y);
  count_launch();
  return e;
}

}  // namespace fdl
