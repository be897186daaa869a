// pw_kernel.cu -- streaming pointwise (1x1) convolution for the large maps: out[M][N] = act(in[M][K] * W[K][N] + bias).
//
// The iris net's bottleneck blocks start with a 1x1 reduction (64 -> 32 at 32x32, 128 -> 64 at 16x16; B = 512 eyes: 0.5 M
// pixels per launch).  These are pure streaming GEMMs with tiny N: 2.1 GFLOP against 200 MB of traffic, i.e. as much FMA-pipe
// time (29 us at the measured 74 TFLOP/s) as HBM time (31 us).  On the tensor-core path every 128-pixel tile is one CTA-long
// latency chain (gather -> hi/lo planes -> MMA -> epilogue, 120 us per launch); here persistent CTAs stream the pixels through
// shared memory with cp.async (double buffered, whole 128-pixel x K tiles, contiguous in NHWC) and do the arithmetic in fp32
// registers with packed FFMA2 -- exact fp32, no operand splitting.
//
//   thread  = 4 pixels (g, g+32, g+64, g+96 of the tile) x 8 output channels: 32 accumulators as 16 packed pairs
//   warp    = one group of 8 output channels for all 128 pixels of the tile; N / 8 warps per CTA
//   shared  = W[K][N] (staged once per CTA) + 2 x [128][K + 4] input tiles (pixel stride = odd number of quads: the per-pixel
//             16-byte reads of a warp are bank-conflict free)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "net_kernels.cuh"
#include "pdl.h"
#include "plan.h"

namespace fdl {

namespace {

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
__device__ __forceinline__ float lo_of(ull v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi_of(ull v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;               // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// PXT pixels per thread (tile = 32 * PXT pixels); two input stages.  (Three stages, and 2 pixels per thread with twice the resident
// CTAs, were measured: 85 / 94 us against 82 us for this configuration on the 64 -> 32 launch.)
template <int PXT>
__global__ void __launch_bounds__(256) pw_stream_kernel(const ConvArgs a, const long long m_max) {
  constexpr int kTilePx = 32 * PXT;
  extern __shared__ __align__(16) float sm[];
  const int K = a.K, N = a.N, KP = K + 4;        // KP: pixel stride of the staged tile (floats)
  float* s_w = sm;                               // [K][N]
  float* s_in0 = sm + K * N;                     // 2 x [kTilePx][KP]
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int lane = tid & 31, cg = tid >> 5, n0 = cg * 8;

  // weights and per-thread constants do not depend on the previous launch (PDL, see pdl.h)
  for (int i = tid; i < K * N / 4; i += nthreads) reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + n0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 4));
  float al[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) al[j] = a.act == ACT_PRELU ? __ldg(a.alpha + n0 + j) : 0.f;
  pdl_launch_dependents();
  pdl_wait();

  long long M = m_max;
  if (a.n_active) M = (long long)min(a.B, *a.n_active) * a.out.H * a.out.W;
  const long long ntiles = (M + kTilePx - 1) / kTilePx;
  if ((long long)blockIdx.x >= ntiles) return;

  const int qpp = K >> 2;                        // 16-byte pieces per pixel
  auto issue_tile = [&](long long tile, int buf) {
    const long long p0 = tile * kTilePx;
    const float* src = a.in.p + p0 * K;          // the tile is contiguous: pixels flattened over the batch
    float* dst = s_in0 + buf * kTilePx * KP;
    for (int c = tid; c < kTilePx * qpp; c += nthreads) {
      const int px = c / qpp, q = c - px * qpp;
      const bool ok = p0 + px < M;
      cp_async16(dst + px * KP + 4 * q, ok ? src + (long long)c * 4 : a.in.p, ok);
    }
    cp_async_commit();
  };
  issue_tile(blockIdx.x, 0);

  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const long long next = tile + gridDim.x;
    if (next < ntiles) { issue_tile(next, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                             // this tile (and, first time, the weights) are visible to every thread
    const float* s_in = s_in0 + buf * kTilePx * KP;

    ull acc[PXT][4];
    {
      const ull p01 = pack2(b0.x, b0.y), p23 = pack2(b0.z, b0.w), p45 = pack2(b1.x, b1.y), p67 = pack2(b1.z, b1.w);
#pragma unroll
      for (int i = 0; i < PXT; ++i) { acc[i][0] = p01; acc[i][1] = p23; acc[i][2] = p45; acc[i][3] = p67; }
    }
    const float* xin = s_in + lane * KP;         // pixels lane, lane+32, ...
    for (int kq = 0; kq < qpp; ++kq) {
      float4 x[PXT];
#pragma unroll
      for (int i = 0; i < PXT; ++i) x[i] = *reinterpret_cast<const float4*>(xin + i * 32 * KP + 4 * kq);
      const float* wp = s_w + (4 * kq) * N + n0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp + kk * N), w1 = *reinterpret_cast<const ulonglong2*>(wp + kk * N + 4);
#pragma unroll
        for (int i = 0; i < PXT; ++i) {
          const float xv = kk == 0 ? x[i].x : (kk == 1 ? x[i].y : (kk == 2 ? x[i].z : x[i].w));
          const ull xx = pack2(xv, xv);
          acc[i][0] = fma2(xx, w0.x, acc[i][0]); acc[i][1] = fma2(xx, w0.y, acc[i][1]);
          acc[i][2] = fma2(xx, w1.x, acc[i][2]); acc[i][3] = fma2(xx, w1.y, acc[i][3]);
        }
      }
    }
    const long long p0 = tile * kTilePx;
#pragma unroll
    for (int i = 0; i < PXT; ++i) {
      const long long m = p0 + lane + 32 * i;
      if (m >= M) continue;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = (j & 1) ? hi_of(acc[i][j >> 1]) : lo_of(acc[i][j >> 1]);
        if (a.act == ACT_RELU) t = fmaxf(t, 0.f);
        else if (a.act == ACT_PRELU) t = t >= 0.f ? t : t * al[j];
        v[j] = t;
      }
      float4* op = reinterpret_cast<float4*>(a.out.p + m * N + n0);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();                             // everyone is done with this buffer before it is refilled
  }
}

size_t pw_smem(int K, int N) { return (size_t)(K * N + 2 * 128 * (K + 4)) * sizeof(float); }

bool pw_enabled() {
  static const bool on = [] { const char* e = getenv("FDL_PW"); return e ? atoi(e) != 0 : true; }();
  return on;
}

}  // namespace

// CONV_2D 1x1, stride 1, no residual, contiguous NHWC in and out, K <= 64 (multiple of 4), N <= 64 (multiple of 8), and enough
// pixels to fill the machine (small maps stay on the tensor-core kernel, where the batch is the only parallelism).
bool pw_stream_supported(const Step& s, int B) {
  if (!pw_enabled() || s.kind != STEP_CONV || s.kh != 1 || s.kw != 1 || s.stride != 1 || s.pad_t != 0 || s.pad_l != 0) return false;
  if (s.skip.tensor >= 0 || s.w < 0) return false;
  const int K = s.in.C, N = s.out.C;
  if (K % 4 != 0 || K > 64 || K < 16 || N % 8 != 0 || N > 64 || s.Npad != N || s.K != K) return false;   // K = 128: no faster than conv_tc (69 us both)
  if (s.in.H != s.out.H || s.in.W != s.out.W) return false;
  if (s.in.offset != 0 || s.out.offset != 0 || s.in.batch_stride != (int64_t)s.in.H * s.in.W * K || s.out.batch_stride != (int64_t)s.out.H * s.out.W * N)
    return false;
  if (pw_smem(K, N) > 200 * 1024) return false;
  return (long long)B * s.out.H * s.out.W >= 148LL * 2 * 128;
}

cudaError_t pw_stream_init() { return cudaFuncSetAttribute(pw_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }

cudaError_t launch_pw_stream(const ConvArgs& a, cudaStream_t stream) {
  const long long M = (long long)a.B * a.out.H * a.out.W;
  if (M <= 0) return cudaSuccess;
  const size_t smem = pw_smem(a.K, a.N);
  const int threads = 32 * (a.N / 8);
  int per_sm = (int)((228 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const long long ntiles = (M + 127) / 128;
  const long long cap = (long long)persist_sms() * per_sm;
  const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
  cudaError_t e = launch_pdl(pw_stream_kernel<4>, dim3(grid), dim3(threads), smem, stream, a, M);
  count_launch();
  return e;
}

}  // namespace fdl
