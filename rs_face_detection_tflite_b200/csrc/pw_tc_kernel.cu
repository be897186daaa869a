// pw_tc_kernel.cu -- large-map 1x1 convolutions (the "reduce" halves of the iris graph's bottleneck blocks: 64 -> 32 at 32 x 32,
// 128 -> 64 at 16 x 16) as a streaming GEMM on the tensor cores, warp-specialised like the stem (stem_tc_kernel.cu).
//
// A dense NHWC tensor is a [pixels][C] matrix: a tile is 128 consecutive pixels (no halo, any position), brought in by TMA in chunks
// of 64 channels; the builders split each chunk into f16 hi / lo planes (the K-major core-matrix layout of tcgen05.mma kind::f16),
// one 16-channel K step at a time, each with its own "full" barrier so that the MMAs of step k run while step k + 1 is laid out;
// the accumulator (128 pixels x N) lives in TMEM; two epilogue teams take alternate tiles (+ bias, PRELU / RELU, staging, one TMA
// store per tile).  The general tensor-core convolution (conv_tc_kernel.cu) gathers its operand with per-thread global loads and
// runs these layers at 0.24 .. 0.41 of their HBM floor (70 / 80 us); this kernel streams them.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "mma_kernels.cuh"
#include "net_kernels.cuh"
#include "pdl.h"
#include "plan.h"
#include "sm100_ptx.cuh"

namespace fdl {

void count_launch();
bool encode_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, long long bstride, int box_h, int box_w, int box_c = 0);
bool encode_tiled4(CUtensorMap* m, const float* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3], const unsigned box[4]);

namespace {

constexpr int kTile = 128;                     // pixels per tile == UMMA M
constexpr int kChunk = 64;                     // channels per TMA chunk: 4 K steps of 16
constexpr int kSteps = kChunk / 16;
constexpr int kEpiThreads = 256, kBuildThreads = 256, kThreads = kEpiThreads + kBuildThreads + 32;
constexpr int kPlane = kTile * 16 + 16;        // one 8-channel plane of A: 128 rows x 16 B (+ 16 B of bank skew)
constexpr int kInStride = kChunk + 4;          // pixel stride of a staged chunk in floats: an odd number of 16-byte quads
constexpr int kMaxSmem = 227 * 1024;

struct TmapQuad { CUtensorMap m[4]; };          // one input view per kernel tap (1x1: only m[0])

struct PwTcArgs {
  const float* w = nullptr;      // [K4][Npad] fp32
  const float* bias = nullptr;
  const float* alpha = nullptr;
  int C = 0, N = 0, Npad = 0, Np = 0, act = 0;
  int NS = 3;                    // chunks in flight
  int teams = 2;                 // epilogue teams (each owns a staging tile)
  int taps = 1;                  // 1: 1x1 convolution over the flattened pixels; 4: 2x2 / stride 2 (K = 4 C, one strided input view per tap)
  int tiles_x = 0, tiles_y = 0, Ho = 0, Wo = 0;   // taps == 4: 8 x 16 output tiles per item
  long long pixels_per_item = 0; // H * W (taps == 1)
  int B = 0;
  const int* n_active = nullptr;
};

struct Layout { int bias, alpha, w, in0, in_stage, a0, out0, out_stage, total; };
__host__ __device__ inline int align_up_p(int v, int a) { return (v + a - 1) / a * a; }
__host__ __device__ inline Layout layout(int K, int N, int Np, int NS, int teams) {
  Layout L;
  int off = 128;                               // barriers + tmem slot
  L.bias = off; off += Np * 4;
  L.alpha = off; off += Np * 4;
  off = align_up_p(off, 128);
  L.w = off; off += 2 * (K / 8) * Np * 16;     // hi planes then lo planes: [K/8][Np][8 halves] each
  off = align_up_p(off, 128);
  L.in_stage = align_up_p(kTile * kInStride * 4, 128);
  L.in0 = off; off += NS * L.in_stage;
  L.a0 = off; off += align_up_p(2 * (kChunk / 8) * kPlane, 128);
  L.out_stage = align_up_p(kTile * (((N >> 2) | 1) << 2) * 4, 128);
  L.out0 = off; off += teams * L.out_stage;
  L.total = align_up_p(off, 128);
  return L;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float c0, float c1) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(c1), "f"(c0));
  return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t d) {
  float2 r;
  asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}\n" : "=f"(r.x), "=f"(r.y) : "r"(d));
  return r;
}
__device__ __forceinline__ uint16_t f2h(float v) {
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return h;
}
__device__ __forceinline__ float h2f(uint16_t h) {
  float v;
  asm("cvt.f32.f16 %0, %1;" : "=f"(v) : "h"(h));
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) pw_tc_kernel(const __grid_constant__ TmapQuad tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                            const PwTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int C = a.C, N = a.N, Np = a.Np, NS = a.NS, cpt = C / kChunk, nchunks = a.taps * cpt, K = a.taps * C;
  const Layout L = layout(K, N, Np, NS, a.teams);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem);       // [NS <= 4]  chunk landed
  uint64_t* a_full = in_full + 4;                              // [kSteps]   the two A planes of K step ks written by every builder
  uint64_t* a_empty = a_full + kSteps;                         //            the MMAs have read the A planes
  uint64_t* acc_full = a_empty + 1;                            // [2]        accumulator complete
  uint64_t* acc_empty = acc_full + 2;                          // [2]        accumulator drained by its epilogue team
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  // ---- prologue: nothing here depends on the previous launch (PDL) ----
  if (tid == 0) {
    for (int i = 0; i < a.taps; ++i) ptx::prefetch_tmap(&tm_in.m[i]);
    ptx::prefetch_tmap(&tm_out);
    for (int s = 0; s < NS; ++s) ptx::mbar_init(&in_full[s], 1);
    for (int k = 0; k < kSteps; ++k) ptx::mbar_init(&a_full[k], kBuildThreads);
    ptx::mbar_init(a_empty, 1);
    for (int t = 0; t < 2; ++t) { ptx::mbar_init(&acc_full[t], 1); ptx::mbar_init(&acc_empty[t], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, (uint32_t)(2 * Np));
  {
    // weights: fp32 [k][n] in global memory -> f16 hi / lo planes [k / 8][n][k % 8]
    uint16_t* w_hi = reinterpret_cast<uint16_t*>(smem + L.w);
    uint16_t* w_lo = w_hi + (K / 8) * Np * 8;
    for (int i = tid; i < (K / 8) * Np * 8; i += kThreads) {
      const int kq = i / (Np * 8), r = i - kq * Np * 8, n = r >> 3, e = r & 7;
      float v = 0.f;
      if (n < N) v = __ldg(a.w + (long long)(8 * kq + e) * a.Npad + n);
      const uint16_t h = f2h(v);
      w_hi[i] = h;
      w_lo[i] = f2h(v - h2f(h));
    }
    for (int i = tid; i < Np; i += kThreads) {
      reinterpret_cast<float*>(smem + L.bias)[i] = i < N ? __ldg(a.bias + i) : 0.f;
      reinterpret_cast<float*>(smem + L.alpha)[i] = (a.alpha && i < N) ? __ldg(a.alpha + i) : 0.f;
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const long long npix = (long long)nb * a.pixels_per_item;
  const int tiles_per_item = a.tiles_x * a.tiles_y;
  const int ntiles = a.taps == 1 ? (int)((npix + kTile - 1) / kTile) : nb * tiles_per_item;
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int my_iters = my_tiles * nchunks;                      // (tile, chunk) pairs, chunk fastest
  const uint32_t in_bytes = (uint32_t)(kTile * kInStride * 4);

  if (my_tiles == 0) {
    // nothing to do
  } else if (warp_u < 4 * a.teams) {
    // ================= epilogue teams: team e takes the CTA's tiles e, e + teams, ... and owns staging buffer e =================
    const int e = warp_u >> 2, p = tid & 127;                  // TMEM lane == pixel of the tile
    const int NPf = ((N >> 2) | 1) << 2;                       // staging pixel stride (floats): an odd number of quads
    float* s_o = reinterpret_cast<float*>(smem + L.out0 + e * L.out_stage) + p * NPf;
    const uint32_t taddr0 = tmem_base + ((uint32_t)((warp_u & 3) * 32) << 16);
    const bool leader = p == 0;
    const int bar_id = 1 + e;
    const float* s_bias = reinterpret_cast<const float*>(smem + L.bias);
    const float* s_alpha = reinterpret_cast<const float*>(smem + L.alpha);
    for (int it = e; it < my_tiles; it += a.teams) {
      const int t = it & 1;                                    // accumulator of the tile
      const uint32_t taddr = taddr0 + (uint32_t)(t * Np);
      ptx::mbar_wait(&acc_full[t], (uint32_t)((it >> 1) & 1));
      ptx::tc_fence_after_sync();
      if (leader) ptx::tma_store_wait_read0();                 // this team's previous store has read the staging buffer
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      for (int c0 = 0; c0 < Np; c0 += 32) {
        uint32_t r0[16], r1[16];
        ptx::tmem_ld16_issue(taddr + (uint32_t)c0, r0);
        ptx::tmem_ld16_issue(taddr + (uint32_t)c0 + 16u, r1);
        ptx::tmem_ld_wait16(r0);
        ptx::tmem_ld_wait16(r1);
        if (c0 + 32 >= Np) {                                   // accumulator drained
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[t]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = c0 + 16 * h + 4 * j;
            if (n < N) {
              const uint32_t* r = h ? r1 : r0;
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n);
              float4 o = make_float4(__uint_as_float(r[4 * j]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y, __uint_as_float(r[4 * j + 2]) + b4.z,
                                     __uint_as_float(r[4 * j + 3]) + b4.w);
              if (a.act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              else if (a.act == ACT_PRELU) {
                const float4 al = *reinterpret_cast<const float4*>(s_alpha + n);
                o.x = o.x >= 0.f ? o.x : o.x * al.x; o.y = o.y >= 0.f ? o.y : o.y * al.y; o.z = o.z >= 0.f ? o.z : o.z * al.z; o.w = o.w >= 0.f ? o.w : o.w * al.w;
              }
              *reinterpret_cast<float4*>(s_o + n) = o;
            }
          }
        }
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        if (a.taps == 1) ptx::tma_store_4d(&tm_out, smem + L.out0 + e * L.out_stage, 0, tile * kTile, 0, 0);
        else {
          const int b = tile / tiles_per_item, r = tile - b * tiles_per_item, ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
          ptx::tma_store_4d(&tm_out, smem + L.out0 + e * L.out_stage, 0, tx * 16, ty * 8, b);
        }
        ptx::tma_store_commit();
      }
    }
    if (leader) ptx::tma_store_wait_all0();
  } else if (warp_u < 8) {
    // (the second epilogue team's warps when only one team runs)
  } else if (warp_u < (kEpiThreads + kBuildThreads) / 32) {
    // ================= builders: chunk -> A planes (hi, lo), one 16-channel K step at a time =================
    const int bt = tid - kEpiThreads, p = bt & 127, half = bt >> 7;     // one pixel, the first or the second 8 channels of a K step
    uint8_t* s_a = smem + L.a0;
    constexpr int KQ = kChunk / 8;                                       // planes per operand half of a chunk
    for (int j = 0; j < my_iters; ++j) {
      const int s = j % NS;
      ptx::mbar_wait(&in_full[s], (uint32_t)((j / NS) & 1));
      if (j > 0) ptx::mbar_wait(a_empty, (uint32_t)((j - 1) & 1));
      const float* src = reinterpret_cast<const float*>(smem + L.in0 + s * L.in_stage) + p * kInStride + 8 * half;
      float4 vv[kSteps][2];
#pragma unroll
      for (int ks = 0; ks < kSteps; ++ks) {
        vv[ks][0] = *reinterpret_cast<const float4*>(src + 16 * ks);
        vv[ks][1] = *reinterpret_cast<const float4*>(src + 16 * ks + 4);
      }
#pragma unroll
      for (int ks = 0; ks < kSteps; ++ks) {
        const int kq = 2 * ks + half;
        const float v[8] = {vv[ks][0].x, vv[ks][0].y, vv[ks][0].z, vv[ks][0].w, vv[ks][1].x, vv[ks][1].y, vv[ks][1].z, vv[ks][1].w};
        uint4 hi, lo;
        hi.x = pack_f16x2(v[0], v[1]); hi.y = pack_f16x2(v[2], v[3]); hi.z = pack_f16x2(v[4], v[5]); hi.w = pack_f16x2(v[6], v[7]);
        const float2 h0 = unpack_f16x2(hi.x), h1 = unpack_f16x2(hi.y), h2 = unpack_f16x2(hi.z), h3 = unpack_f16x2(hi.w);
        lo.x = pack_f16x2(v[0] - h0.x, v[1] - h0.y); lo.y = pack_f16x2(v[2] - h1.x, v[3] - h1.y);
        lo.z = pack_f16x2(v[4] - h2.x, v[5] - h2.y); lo.w = pack_f16x2(v[6] - h3.x, v[7] - h3.y);
        *reinterpret_cast<uint4*>(s_a + kq * kPlane + p * 16) = hi;
        *reinterpret_cast<uint4*>(s_a + (KQ + kq) * kPlane + p * 16) = lo;
        ptx::fence_proxy_async_smem();
        mbar_arrive(&a_full[ks]);
      }
    }
  } else {
    // ================= the last warp: input ring + MMA issue (the whole warp, converged; one elected lane issues) =================
    auto issue_load = [&](int j) {
      const int it = j / nchunks, c = j - it * nchunks, s = j % NS;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      ptx::mbar_arrive_expect_tx(&in_full[s], in_bytes);
      if (a.taps == 1) ptx::tma_load_4d(smem + L.in0 + s * L.in_stage, &tm_in.m[0], &in_full[s], c * kChunk, tile * kTile, 0, 0);
      else {
        // chunk c = (tap, 64-channel group); the tap's view is the input sampled at (2 y + dy, 2 x + dx): an 8 x 16 tile of output pixels
        const int tap = c / cpt, cc = c - tap * cpt;
        const int b = tile / tiles_per_item, r = tile - b * tiles_per_item, ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
        ptx::tma_load_4d(smem + L.in0 + s * L.in_stage, &tm_in.m[tap], &in_full[s], cc * kChunk, tx * 16, ty * 8, b);
      }
    };
    if (lane == 0)
      for (int j = 0; j < NS && j < my_iters; ++j) issue_load(j);
    __syncwarp();
    const uint32_t idesc = ptx::umma_idesc_f16(128, Np);
    constexpr int KQ = kChunk / 8;
    const uint32_t a_hi = ptx::smem_u32(smem + L.a0), a_lo = a_hi + (uint32_t)(KQ * kPlane);
    const uint32_t w_hi = ptx::smem_u32(smem + L.w), w_lo = w_hi + (uint32_t)((K / 8) * Np * 16);
    const uint32_t lbo_w = (uint32_t)Np * 16u;
    for (int j = 0; j < my_iters; ++j) {
      const int it = j / nchunks, c = j - it * nchunks, t = it & 1;
      const uint32_t d_tmem = tmem_base + (uint32_t)(t * Np);
#pragma unroll
      for (int ks = 0; ks < kSteps; ++ks) {
        ptx::mbar_wait(&a_full[ks], (uint32_t)(j & 1));
        if (ks == 0 && c == 0 && it >= 2) ptx::mbar_wait(&acc_empty[t], (uint32_t)(((it >> 1) - 1) & 1));
        if (ks == kSteps - 1) {
          // every builder is done with this chunk: its stage takes the chunk NS places ahead
          if (lane == 0 && j + NS < my_iters) issue_load(j + NS);
          __syncwarp();
        }
        ptx::tc_fence_after_sync();
        const uint32_t wq = (uint32_t)(c * KQ + 2 * ks);       // first weight plane of this K step
        const uint64_t dah = ptx::umma_desc_kmajor(a_hi + (uint32_t)(2 * ks * kPlane), kPlane, 128);
        const uint64_t dal = ptx::umma_desc_kmajor(a_lo + (uint32_t)(2 * ks * kPlane), kPlane, 128);
        const uint64_t dbh = ptx::umma_desc_kmajor(w_hi + wq * lbo_w, lbo_w, 128);
        const uint64_t dbl = ptx::umma_desc_kmajor(w_lo + wq * lbo_w, lbo_w, 128);
        ptx::mma_f16_elect(d_tmem, dah, dbh, idesc, (c | ks) ? 1u : 0u);
        ptx::mma_f16_elect(d_tmem, dal, dbh, idesc, 1u);
        ptx::mma_f16_elect(d_tmem, dah, dbl, idesc, 1u);
      }
      ptx::mma_commit_elect(a_empty);
      if (c == nchunks - 1) ptx::mma_commit_elect(&acc_full[t]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)(2 * Np));
}

// (NS, teams) for a contraction of K values: as many chunks in flight as fit, two epilogue teams if they still leave two stages
bool pick_cfg(int K, int N, int Np, int* ns, int* teams) {
  for (int tm = 2; tm >= 1; --tm)
    for (int s = 4; s >= 2; --s)
      if (layout(K, N, Np, s, tm).total <= kMaxSmem) { *ns = s; *teams = tm; return true; }
  return false;
}

}  // namespace

cudaError_t pw_tc_init() { return cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); }

// Output channels per launch: all of them, or halves when the weights of the whole contraction do not fit next to the pipeline
static int pw_tc_split(int K, int N) {
  int ns, tm;
  for (int parts = 1; parts <= 2; ++parts) {
    const int n = N / parts;
    if (N % parts || n % 16) continue;
    if (pick_cfg(K, n, n <= 32 ? 32 : 64, &ns, &tm)) return parts;
  }
  return 0;
}

bool pw_tc_supported(const Step& s, int B) {
  static const bool on = [] { const char* e = getenv("FDL_PW_TC"); return e ? atoi(e) != 0 : true; }();
  (void)B;   // never a function of the batch: a frame's result must not depend on how many frames travel with it
  if (!on || s.kind != STEP_CONV || s.pad_t != 0 || s.pad_l != 0 || s.skip.tensor >= 0 || s.kh != s.kw) return false;
  const int C = s.in.C, N = s.out.C;
  if (C % kChunk != 0 || C > 128 || N % 16 != 0 || N > 64 || N < 16) return false;
  if (s.in.offset != 0 || s.out.offset != 0 || s.in.batch_stride != (int64_t)s.in.H * s.in.W * C || s.out.batch_stride != (int64_t)s.out.H * s.out.W * N) return false;
  if (s.kh == 1 && s.stride == 1) {
    if (s.in.H != s.out.H || s.in.W != s.out.W) return false;
    if (s.in.H * s.in.W < 128) return false;                                 // tiny maps: the general kernel (or the tail chain)
    return pw_tc_split(C, N) == 1;
  }
  if (s.kh == 2 && s.stride == 2) {                                          // 2x2 / stride 2 (the iris net's down-sampling convolutions)
    static const bool patch_on = [] { const char* e = getenv("FDL_PW_TC_PATCH"); return e ? atoi(e) != 0 : true; }();
    if (!patch_on || s.in.H != 2 * s.out.H || s.in.W != 2 * s.out.W || s.K != 4 * C) return false;
    if (s.out.W % 16 != 0 || s.out.H % 8 != 0) return false;               // whole 8 x 16 output tiles (half-empty tiles and a split over
    return pw_tc_split(4 * C, N) == 1;                                       // the output channels measured slower: 84 vs 57 us at 8 x 8)
  }
  return false;
}

cudaError_t launch_pw_tc(const ConvArgs& a, cudaStream_t stream) {
  const int taps = a.kh * a.kw, C = a.in.C, K = taps * C;
  const int parts = pw_tc_split(K, a.N);
  if (parts == 0) return cudaErrorInvalidConfiguration;
  const int Nl = a.N / parts;
  PwTcArgs k;
  k.C = C; k.N = Nl; k.Npad = a.Npad; k.Np = Nl <= 32 ? 32 : 64; k.act = a.act; k.taps = taps;
  if (!pick_cfg(K, Nl, k.Np, &k.NS, &k.teams)) return cudaErrorInvalidConfiguration;
  k.pixels_per_item = (long long)a.in.H * a.in.W; k.B = a.B; k.n_active = a.n_active;
  const long long npix = (long long)a.B * k.pixels_per_item;
  if (npix > 0x7fffff00LL) return cudaErrorInvalidValue;      // tile indices and TMA coordinates are 32-bit
  const Layout L = layout(K, Nl, k.Np, k.NS, k.teams);
  int ntiles;
  if (taps == 1) ntiles = (int)((npix + kTile - 1) / kTile);
  else {
    k.Ho = a.out.H; k.Wo = a.out.W; k.tiles_x = (a.out.W + 15) / 16; k.tiles_y = (a.out.H + 7) / 8;
    ntiles = a.B * k.tiles_x * k.tiles_y;
  }
  if (ntiles == 0) return cudaSuccess;
  int grid = persist_sms();
  if (grid > ntiles) grid = ntiles;
  for (int part = 0; part < parts; ++part) {
    const int n0 = part * Nl;
    k.w = a.w + n0; k.bias = a.bias + n0; k.alpha = a.alpha ? a.alpha + n0 : nullptr;
    TmapQuad tm_in;
    CUtensorMap tm_out;
    if (taps == 1) {
      // the tensors as [pixels][C] matrices: dims {C, pixels, 1, 1}, boxes {C chunk (+ 4 floats of padding), 128 pixels}
      if (!encode_nhwc(&tm_in.m[0], a.in.p, 1, 1, (int)npix, C, npix * C, 1, kTile, kInStride)) return cudaErrorInvalidValue;
      tm_in.m[1] = tm_in.m[2] = tm_in.m[3] = tm_in.m[0];
      if (!encode_nhwc(&tm_out, a.out.p, 1, 1, (int)npix, Nl, npix * Nl, 1, kTile, ((Nl / 4) | 1) * 4)) return cudaErrorInvalidValue;
    } else {
      // tap (dy, dx): the input sampled at every second pixel from (dy, dx) on: dims {C, W/2, H/2, B} with doubled strides
      const unsigned long long W = (unsigned long long)a.in.W;
      const unsigned long long dims[4] = {(unsigned long long)C, (unsigned long long)a.out.W, (unsigned long long)a.out.H, (unsigned long long)a.B};
      const unsigned long long strides[3] = {2ull * C * 4, 2ull * W * C * 4, (unsigned long long)a.in.bstride * 4};
      const unsigned box[4] = {(unsigned)kInStride, 16u, 8u, 1u};
      for (int tap = 0; tap < 4; ++tap)
        if (!encode_tiled4(&tm_in.m[tap], a.in.p + ((long long)(tap >> 1) * a.in.W + (tap & 1)) * C, dims, strides, box)) return cudaErrorInvalidValue;
      // the launch's output channels [n0, n0 + Nl) of the NHWC output
      const unsigned long long odims[4] = {(unsigned long long)Nl, (unsigned long long)a.out.W, (unsigned long long)a.out.H, (unsigned long long)a.B};
      const unsigned long long ostrides[3] = {(unsigned long long)a.N * 4, (unsigned long long)a.out.W * a.N * 4, (unsigned long long)a.out.bstride * 4};
      const unsigned obox[4] = {(unsigned)(((Nl / 4) | 1) * 4), 16u, 8u, 1u};
      if (!encode_tiled4(&tm_out, a.out.p + n0, odims, ostrides, obox)) return cudaErrorInvalidValue;
    }
    cudaError_t e = launch_pdl(pw_tc_kernel, dim3(grid), dim3(kThreads), (size_t)L.total, stream, tm_in, tm_out, k);
    count_launch();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace fdl
