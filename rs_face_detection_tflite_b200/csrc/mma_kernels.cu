// mma_kernels.cu -- the tensor-core BlazeBlock kernel (sm_100a: TMA + tcgen05 + TMEM).
//
// One launch computes, for a batch of NHWC f32 feature maps,
//     out = act( PW1x1( DW3x3_s(in) + b_dw ) + b_pw + skip )          s = 1 or 2
// (SURVEY.md A.2 "Single BlazeBlock" and the two halves of the double / bottleneck blocks) with
//   * the input halo tile ((TH-1)s+3) x ((TW-1)s+3) x C brought in by ONE TMA tensor load per tile
//     (out-of-bounds coordinates are zero-filled by the TMA unit == TFLite SAME padding), double-buffered on
//     mbarriers when shared memory allows,
//   * the depthwise 3x3 on the CUDA cores straight out of shared memory (sliding row window in registers,
//     float4 channel quads), written as the A operand of the pointwise GEMM in the UMMA K-major
//     core-matrix layout,
//   * the pointwise 1x1 contraction [128 pixels x Cin] x [Cin x Cout] on the 5th-gen tensor cores:
//     tcgen05.mma kind::tf32, M=128, accumulator in TMEM.  fp32 fidelity is kept by operand splitting:
//     x = hi + lo with hi = tf32(x), lo = x - hi, so A*W = A_hi*W + A_lo*W (+ A_hi*W_lo when the weights are
//     not tf32-exact; the detectors' f16-stored weights are) -- error ~2^-22 per product, same order as fp32,
//   * epilogue out of TMEM (tcgen05.ld): + bias, + residual (identity or MAX_POOL 2x2 from the resident
//     input tile, or direct / MAX_POOL 2x2 from global; zero channel PAD), RELU / PRELU, staged in shared
//     memory and written with ONE TMA tensor store per tile.
// The depthwise result and the pointwise accumulator never touch HBM: per tile the kernel reads the
// input tile once and writes the output tile once (the "block-fused floor" of SURVEY.md 8d).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "mma_kernels.cuh"
#include "pdl.h"
#include "plan.h"
#include "sm100_ptx.cuh"

namespace fdl {

void count_launch();
bool encode_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, long long bstride, int box_h, int box_w, int box_c = 0);

namespace {

constexpr int TH = 8, TW = 16;            // output tile: 128 pixels == UMMA M
// CTA size: 6 warps.  A 12-warp variant for the one-CTA-per-SM configurations is kept for A/B timing (FDL_TC_THREADS=384):
// measured no faster on B200 (the tile's latency chain, not the depthwise issue rate, bounds this serial kernel).
// Warps 0..3 own the 128 TMEM lanes in the epilogue.
constexpr int kThreadsSmall = 192, kThreadsBig = 384;
constexpr int kPlanePad = 16;             // bytes added to each 2 KB channel-quad plane of A (bank spreading)
constexpr int kPlaneBytes = TH * TW * 16 + kPlanePad;   // LBO of the A operand

struct SmemLayout {
  int bias, alpha, in0, in_stage, a_hi, a_lo, w, out, total;
};

__host__ __device__ inline int align_up_i(int v, int a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int in_tile_h(int S) { return (TH - 1) * S + 3; }
__host__ __device__ inline int in_tile_w(int S) { return (TW - 1) * S + 3; }

// Pixel strides of the input and the output staging tile in shared memory: an ODD number of 16-byte quads, so that the
// epilogue's per-pixel 16-byte accesses (thread == pixel) spread over all banks instead of colliding 8-fold when the channel
// count is a multiple of 32.  The TMA boxes are simply wider than the tensors: loads zero-fill the tail, stores clip it.
// (Only when the quad count is a multiple of 4: that is where the collisions are 4- and 8-fold; for 6 or 7 quads the plain
// stride is already within 2x of conflict-free and keeps the depthwise loads, which walk quads first, perfectly linear.)
__host__ __device__ inline int pad_quads(int c) { return ((c >> 2) & 3) == 0 ? c + 4 : c; }

// f16: the A operand is ONE plane set of (f16 hi, f16 lo) pairs (see BlockTcArgs::f16) and `wsplit` counts f16 weight copies.
__host__ __device__ inline SmemLayout smem_layout(int C, int N, int Np, int S, int stages, int wsplit, int alias_out, int CP, int NP, int f16 = 0) {
  SmemLayout L;
  int off = 64;                                   // barriers + tmem pointer
  L.bias = off; off += Np * 4;
  L.alpha = off; off += Np * 4;
  off = align_up_i(off, 128);
  L.in_stage = align_up_i(in_tile_h(S) * in_tile_w(S) * CP * 4, 128);
  L.in0 = off; off += stages * L.in_stage;
  L.a_hi = off; off += (C / 4) * kPlaneBytes;
  L.a_lo = off; if (!f16) off += (C / 4) * kPlaneBytes;     // contiguous with a_hi (kPlaneBytes is a multiple of 16); absent in f16 mode
  off = align_up_i(off, 128);
  L.w = off; off += wsplit * (C / 4) * Np * 16;
  off = align_up_i(off, 128);
  if (alias_out) { L.out = L.a_hi; }              // the output tile reuses the A planes (dead once the MMA has completed)
  else { L.out = off; off += TH * TW * NP * 4; }
  L.total = align_up_i(off, 128);
  return L;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
  a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// (c0, c1) -> f16x2 with c0 in the low half, round to nearest even, saturating instead of overflowing; and back
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

template <int S, int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 192 ? 2 : 1) blaze_block_tc_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                  const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_skip,
                                                                  const BlockTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ITH = (TH - 1) * S + 3, ITW = (TW - 1) * S + 3;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = a.C, N = a.N, Np = a.Np, Q = C >> 2;
  const bool f16 = a.f16 != 0;
  const int wcopies = f16 ? a.wsplit16 : a.wsplit;
  const SmemLayout L = smem_layout(C, N, Np, S, a.stages, wcopies, a.alias_out, a.tc_cp, a.tc_np, a.f16);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);          // [2]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + 16);
  uint64_t* skip_bar = reinterpret_cast<uint64_t*>(smem + 24);     // the residual tile landed in the output staging tile (skip_tma)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_alpha = reinterpret_cast<float*>(smem + L.alpha);
  uint8_t* s_ahi = smem + L.a_hi;
  uint8_t* s_alo = smem + L.a_lo;
  float* s_w = reinterpret_cast<float*>(smem + L.w);
  float* s_out = reinterpret_cast<float*>(smem + L.out);

  // ---- one-time setup: nothing here depends on the previous launch (PDL, see pdl.h) ----
  if (tid == 0) {
    ptx::prefetch_tmap(&tm_in);
    ptx::prefetch_tmap(&tm_out);
    ptx::mbar_init(&full_bar[0], 1);
    ptx::mbar_init(&full_bar[1], 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::mbar_init(skip_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  for (int i = tid; i < Np; i += kThreads) {
    s_bias[i] = i < N ? a.bias[i] : 0.f;
    s_alpha[i] = (a.alpha && i < N) ? a.alpha[i] : 0.f;
  }
  {  // pointwise weights: already in UMMA core-matrix order in global memory -> straight copy
    const int n4 = wcopies * Q * Np;   // float4 count
    const float4* src = reinterpret_cast<const float4*>(f16 ? a.w_f16 : a.w_umma);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < n4; i += kThreads) dst[i] = __ldg(src + i);
  }
  // depthwise weights of this thread's channel quad (kThreads % Q == 0, so the quad is fixed per thread)
  const int q = tid % Q;
  float4 wd[9], bd;
#pragma unroll
  for (int k = 0; k < 9; ++k) wd[k] = __ldg(reinterpret_cast<const float4*>(a.w_dw + k * C) + q);
  bd = __ldg(reinterpret_cast<const float4*>(a.b_dw) + q);

  ptx::fence_proxy_async_smem();   // s_w was written through the generic proxy, the MMA reads it through the async proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_launch_dependents();
  pdl_wait();                                       // the previous launch's activations (and *n_active) are visible from here on
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int ntiles = nb * tiles_per_img;

  // Stride-2 blocks whose residual is the MAX_POOL 2x2 of their own input: the depthwise threads hold exactly those pixels, so they
  // leave the pooled residual in the output staging tile; nobody reads the input tile after the depthwise then.
  const bool pool_in_dw = S == 2 && a.skip_mode == 4 && !a.alias_out;
  const bool early_load = a.stages == 1 && (a.skip_mode == 0 || a.skip_mode == 2 || pool_in_dw);   // (with the pooled global residual, mode 3, it measured slower: 138 -> 147 us)
  const bool skip_tma = a.skip_tma != 0;
  const int CP = a.tc_cp, NPf = a.tc_np;            // pixel strides (floats) of the input / output staging tiles
  const uint32_t in_bytes = (uint32_t)(ITH * ITW * CP * 4);
  auto issue_load = [&](int tile, int stage) {
    int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    int ty = r / a.tiles_x, tx = r - ty * a.tiles_x;
    ptx::mbar_arrive_expect_tx(&full_bar[stage], in_bytes);
    ptx::tma_load_4d(smem + L.in0 + stage * L.in_stage, &tm_in, &full_bar[stage], 0, tx * TW * S - a.pad, ty * TH * S - a.pad, b);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue_load(blockIdx.x, 0);

  const uint32_t idesc = f16 ? ptx::umma_idesc_f16(128, Np) : ptx::umma_idesc_tf32(128, Np);
  const uint32_t ahi_addr = ptx::smem_u32(s_ahi), alo_addr = ptx::smem_u32(s_alo), w_addr = ptx::smem_u32(s_w);
  const uint32_t w_lbo = (uint32_t)Np * 16u;
  const int nitems = Q * 2 * TW;

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = a.stages == 2 ? (it & 1) : 0;
    const uint32_t full_parity = a.stages == 2 ? ((it >> 1) & 1) : (it & 1);
    const int next = tile + gridDim.x;
    if (a.stages == 2 && tid == 0 && next < ntiles) issue_load(next, stage ^ 1);
    if (a.alias_out) {
      // the previous tile's TMA store reads the region the depthwise is about to overwrite
      if (tid == 0) ptx::tma_store_wait_read0();
      __syncthreads();
    }
    if (skip_tma && tid == 0) {
      // the residual tile (another tensor: the iris bottlenecks) travels by TMA into the output staging tile -- same pixel stride,
      // missing channels zero-filled -- while the depthwise and the MMAs run; the epilogue then updates the tile in place
      ptx::tma_store_wait_read0();                    // the previous tile's store has read the staging tile
      ptx::mbar_arrive_expect_tx(skip_bar, (uint32_t)(TH * TW * NPf * 4));
      int b = tile / tiles_per_img, rr = tile - b * tiles_per_img, ty = rr / a.tiles_x, tx = rr - ty * a.tiles_x;
      ptx::tma_load_4d(s_out, &tm_skip, skip_bar, 0, tx * TW, ty * TH, b);
    }
    if (pool_in_dw) {
      if (tid == 0) ptx::tma_store_wait_read0();      // the previous tile's store has read the staging tile the depthwise writes into
      __syncthreads();
    }
    ptx::mbar_wait(&full_bar[stage], full_parity);
    const float* s_in = reinterpret_cast<const float*>(smem + L.in0 + stage * L.in_stage);

    // ---- depthwise 3x3 -> A operand (hi / lo planes) ----
    for (int item = tid; item < nitems; item += kThreads) {
      const int xr = item / Q;           // item % Q == q
      const int x = xr % TW, half = xr / TW;
      float4 acc[4], pool[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = bd;
      const float* base = s_in + ((half * 4 * S) * ITW + x * S) * CP + 4 * q;
#pragma unroll
      for (int r = 0; r < 3 * S + 3; ++r) {          // input rows feeding 4 consecutive output rows
        const float* rp = base + r * ITW * CP;
        float4 v0 = ld4(rp), v1 = ld4(rp + CP), v2 = ld4(rp + 2 * CP);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int ky = r - o * S;
          if (ky >= 0 && ky < 3) {
            fma4(acc[o], v0, wd[ky * 3 + 0]);
            fma4(acc[o], v1, wd[ky * 3 + 1]);
            fma4(acc[o], v2, wd[ky * 3 + 2]);
          }
          if (S == 2 && ky == 0) pool[o] = max4(v0, v1);                     // the 2x2 window of output (o, x): rows 2o, 2o + 1,
          if (S == 2 && ky == 1) pool[o] = max4(pool[o], max4(v0, v1));      // columns 2x, 2x + 1 (SAME pad 0 before)
        }
      }
      if (pool_in_dw && 4 * q < a.skip_c) {
#pragma unroll
        for (int o = 0; o < 4; ++o) *reinterpret_cast<float4*>(s_out + ((half * 4 + o) * TW + x) * NPf + 4 * q) = pool[o];
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int p = (half * 4 + o) * TW + x;
        float4 v = acc[o], hi, lo;
        if (f16) {
          // (f16 hi, f16 lo) of the four channels in ONE 16-byte core-matrix row: K' = (hi c0..c3, lo c0..c3)
          uint4 u;
          u.x = pack_f16x2(v.x, v.y); u.y = pack_f16x2(v.z, v.w);
          const float2 h01 = unpack_f16x2(u.x), h23 = unpack_f16x2(u.y);
          u.z = pack_f16x2(v.x - h01.x, v.y - h01.y); u.w = pack_f16x2(v.z - h23.x, v.w - h23.y);
          *reinterpret_cast<uint4*>(s_ahi + q * kPlaneBytes + p * 16) = u;
          continue;
        }
        hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        *reinterpret_cast<float4*>(s_ahi + q * kPlaneBytes + p * 16) = hi;
        *reinterpret_cast<float4*>(s_alo + q * kPlaneBytes + p * 16) = lo;
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    // one input buffer and nobody reads it after the depthwise (no residual from the resident tile): the next tile's load starts now,
    // under the MMAs and the epilogue, not after them
    if (early_load && tid == 0 && next < ntiles) issue_load(next, 0);

    // ---- pointwise 1x1 on the tensor cores: warp 0 (converged) runs the issue loop, one elected lane issues -- the operands stay on
    // the uniform datapath (see mma_f16_elect) -- completion arrives on mma_bar ----
    if (warp_u == 0) {
      ptx::tc_fence_after_sync();
      uint32_t acc_flag = 0;
      if (f16) {
        // kind::f16, K = 16 = two planes = two channel quads x (hi, lo); pass 0: (A_hi, A_lo) * (W, W), pass 1: (A_hi, A_lo) * (W_lo, 0)
        for (int pass = 0; pass < a.wsplit16; ++pass) {
          const uint32_t b_base = w_addr + (uint32_t)(pass * Q * Np * 16);
          for (int ks = 0; ks < (Q >> 1); ++ks) {
            uint64_t ad = ptx::umma_desc_kmajor(ahi_addr + (uint32_t)(ks * 2 * kPlaneBytes), kPlaneBytes, 128);
            uint64_t bdsc = ptx::umma_desc_kmajor(b_base + (uint32_t)ks * 2u * w_lbo, w_lbo, 128);
            ptx::mma_f16_elect(tmem_base, ad, bdsc, idesc, acc_flag);
            acc_flag = 1;
          }
        }
      }
      const int ksteps = C >> 3;
      for (int pass = 0; pass < (a.wsplit == 2 ? 3 : 2) && !f16; ++pass) {
        // pass 0: A_lo * W_hi, pass 1: A_hi * W_hi, pass 2: A_hi * W_lo
        const uint32_t a_base = pass == 0 ? alo_addr : ahi_addr;
        const uint32_t b_base = pass == 2 ? w_addr + (uint32_t)(Q * Np * 16) : w_addr;
        for (int ks = 0; ks < ksteps; ++ks) {
          uint64_t ad = ptx::umma_desc_kmajor(a_base + (uint32_t)(ks * 2 * kPlaneBytes), kPlaneBytes, 128);
          uint64_t bdsc = ptx::umma_desc_kmajor(b_base + (uint32_t)ks * 2u * w_lbo, w_lbo, 128);
          ptx::mma_tf32_elect(tmem_base, ad, bdsc, idesc, acc_flag);
          acc_flag = 1;
        }
      }
      ptx::mma_commit_elect(mma_bar);
    }

    // ---- epilogue: TMEM -> registers -> (+bias, +skip, act) -> smem -> TMA store ----
    if (warp < 4) {
      ptx::mbar_wait(mma_bar, (uint32_t)(it & 1));
      ptx::tc_fence_after_sync();
      if (skip_tma) ptx::mbar_wait(skip_bar, (uint32_t)(it & 1));
      else if (pool_in_dw) { }                       // (the staging tile was handed over before the depthwise wrote the residual into it)
      else if (!a.alias_out) {
        if (tid == 0) ptx::tma_store_wait_read0();   // the previous tile's store has finished reading s_out
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int p = tid;                             // TMEM lane == pixel in tile
      const int py = p / TW, px = p - py * TW;
      int b = tile / tiles_per_img, rr = tile - b * tiles_per_img;
      int ty = rr / a.tiles_x, tx = rr - ty * a.tiles_x;
      const int oy = ty * TH + py, ox = tx * TW + px;
      const bool inside = oy < a.H && ox < a.W;
      // residual sources inside the resident input tile: centre pixel (stride 1) / 2x2 window (stride 2, pad 0)
      const float* skip_smem = S == 1 ? s_in + ((py + 1) * ITW + (px + 1)) * CP : s_in + ((2 * py) * ITW + 2 * px) * CP;
      const float* skip_g = nullptr;
      if ((a.skip_mode == 2 || a.skip_mode == 3) && inside) {
        if (a.skip_mode == 2) skip_g = a.skip + (long long)b * a.skip_bstride + ((long long)oy * a.W + ox) * a.skip_c;
        else skip_g = a.skip + (long long)b * a.skip_bstride + ((long long)(2 * oy) * (2 * a.W) + 2 * ox) * a.skip_c;
      }
      if (a.skip_mode == 2 || a.skip_mode == 3) {
        // Residuals that live in global memory (the bottleneck blocks of the iris net: skip != block input) are fetched 64
        // channels at a time, ALL loads issued before the accumulator columns are read: one global round trip per batch
        // instead of one per channel quad (the ncu source view had the epilogue's FADDs waiting on these loads one by one).
        const long long rs = (long long)2 * a.W * a.skip_c;   // row stride of the MAX_POOL source (skip_mode 3)
        for (int b0 = 0; b0 < Np; b0 += 64) {
          float4 sk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int n = b0 + 4 * j;
            if (skip_tma) { if (n < N) sk[j] = ld4(s_out + p * NPf + n); }
            else if (skip_g && a.skip_mode == 2 && n < a.skip_c) sk[j] = __ldg(reinterpret_cast<const float4*>(skip_g + n));
          }
#pragma unroll
          for (int cc = 0; cc < 64; cc += 16) {
            const int c0 = b0 + cc;
            if (c0 >= Np) break;
            if (skip_g && a.skip_mode == 3) {     // MAX_POOL 2x2 source: four loads per quad -> 16 channels (16 loads) per round trip
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = c0 + 4 * j;
                if (n < a.skip_c)
                  sk[(cc >> 2) + j] = max4(max4(__ldg(reinterpret_cast<const float4*>(skip_g + n)), __ldg(reinterpret_cast<const float4*>(skip_g + a.skip_c + n))),
                                           max4(__ldg(reinterpret_cast<const float4*>(skip_g + rs + n)), __ldg(reinterpret_cast<const float4*>(skip_g + rs + a.skip_c + n))));
              }
            }
            float v[16];
            ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int n = c0 + j;
              if (n >= N) break;
              const float4 s = sk[(cc + j) >> 2];
              float o4[4] = {v[j] + s_bias[n] + s.x, v[j + 1] + s_bias[n + 1] + s.y, v[j + 2] + s_bias[n + 2] + s.z, v[j + 3] + s_bias[n + 3] + s.w};
              if (a.act == ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o4[e] = fmaxf(o4[e], 0.f);
              } else if (a.act == ACT_PRELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o4[e] = o4[e] >= 0.f ? o4[e] : o4[e] * s_alpha[n + e];
              }
              *reinterpret_cast<float4*>(s_out + p * NPf + n) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            }
          }
        }
      } else {
        // residual inside the resident input tile (or none)
        for (int c0 = 0; c0 < Np; c0 += 16) {
          float v[16];
          ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int n = c0 + j;
            if (n >= N) break;
            float o4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] = v[j + e] + s_bias[n + e];
            if (n < a.skip_c) {
              if (a.skip_mode == 1) {
                float4 s = ld4(skip_smem + n);
                o4[0] += s.x; o4[1] += s.y; o4[2] += s.z; o4[3] += s.w;
              } else if (pool_in_dw) {
                const float4 s = ld4(s_out + p * NPf + n);
                o4[0] += s.x; o4[1] += s.y; o4[2] += s.z; o4[3] += s.w;
              } else if (a.skip_mode == 4) {
                float4 s = max4(max4(ld4(skip_smem + n), ld4(skip_smem + CP + n)), max4(ld4(skip_smem + ITW * CP + n), ld4(skip_smem + (ITW + 1) * CP + n)));
                o4[0] += s.x; o4[1] += s.y; o4[2] += s.z; o4[3] += s.w;
              }
            }
            if (a.act == ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o4[e] = fmaxf(o4[e], 0.f);
            } else if (a.act == ACT_PRELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o4[e] = o4[e] >= 0.f ? o4[e] : o4[e] * s_alpha[n + e];
            }
            *reinterpret_cast<float4*>(s_out + p * NPf + n) = make_float4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        ptx::tma_store_4d(&tm_out, s_out, 0, tx * TW, ty * TH, b);
        ptx::tma_store_commit();
      }
    }
    __syncthreads();   // input stage, A planes and the TMEM accumulator are free again
    if (a.stages == 1 && !early_load && tid == 0 && next < ntiles) issue_load(next, 0);
  }

  if (tid == 0) ptx::tma_store_wait_all0();
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;
constexpr int kMaxSmemTc = 227 * 1024;

}  // namespace

bool encode_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, long long bstride, int box_h, int box_w, int box_c) {
  if (!g_encode) return false;
  if (box_c <= 0) box_c = C;   // box_c > C: the tile's pixel stride in shared memory is padded (loads zero-fill, stores clip)
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)bstride * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// A 4-D tiled map over an arbitrary strided view (innermost dimension contiguous; strides in bytes, multiples of 16).
bool encode_tiled4(CUtensorMap* m, const float* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3], const unsigned box[4]) {
  if (!g_encode) return false;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

// f16-split A operand in the serial kernel too (FDL_TC_F16, default on): half the A-plane bytes -> room for a second input stage /
// a second resident CTA for the wide blocks.
bool tc_f16_enabled() {
  static const bool on = [] { const char* e = getenv("FDL_TC_F16"); return e ? atoi(e) != 0 : true; }();
  return on;
}

// Picks (stages, alias_out, pixel strides) for a block; returns false when it cannot fit in shared memory.  The padded
// pixel strides are used only where they cost neither a pipeline stage nor a resident CTA.
bool pick_smem(int C, int N, int Np, int S, int wsplit, int f16, int* stages, int* alias_out, int* total, int* cp_out, int* np_out) {
  // candidate configurations in order of preference at equal occupancy
  const int cand[4][2] = {{2, 0}, {2, 1}, {1, 0}, {1, 1}};
  auto best_for = [&](int CP, int NP, int* per_sm_out) {
    const bool can_alias = TH * TW * NP * 4 <= (f16 ? 1 : 2) * (C / 4) * kPlaneBytes;
    int best = -1, best_per_sm = 0;
    for (int i = 0; i < 4; ++i) {
      if (cand[i][1] && !can_alias) continue;
      SmemLayout L = smem_layout(C, N, Np, S, cand[i][0], wsplit, cand[i][1], CP, NP, f16);
      if (L.total > kMaxSmemTc) continue;
      int per_sm = (228 * 1024) / (L.total + 1024);
      if (per_sm > 2) per_sm = 2;
      if (per_sm > best_per_sm) { best_per_sm = per_sm; best = i; }
    }
    *per_sm_out = best_per_sm;
    return best;
  };
  int ps0 = 0;
  const int plain = best_for(C, N, &ps0);
  if (plain < 0) return false;
  int best = plain, CP = C, NP = N;
  const int opts[3][2] = {{pad_quads(C), pad_quads(N)}, {pad_quads(C), N}, {C, pad_quads(N)}};
  for (int o = 0; o < 3; ++o) {
    if (opts[o][0] == C && opts[o][1] == N) continue;
    int ps = 0;
    const int b = best_for(opts[o][0], opts[o][1], &ps);
    if (b >= 0 && ps >= ps0 && cand[b][0] >= cand[plain][0]) { best = b; CP = opts[o][0]; NP = opts[o][1]; break; }
  }
  *stages = cand[best][0]; *alias_out = cand[best][1]; *cp_out = CP; *np_out = NP;
  *total = smem_layout(C, N, Np, S, *stages, wsplit, *alias_out, CP, NP, f16).total;
  return true;
}

}  // namespace

cudaError_t mma_kernels_init() {
  cudaError_t err = cudaSuccess;
  std::call_once(g_encode_once, [&]() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (err == cudaSuccess && q == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(blaze_block_tc_kernel<1, kThreadsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTc);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(blaze_block_tc_kernel<2, kThreadsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTc);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(blaze_block_tc_kernel<1, kThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTc);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(blaze_block_tc_kernel<2, kThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTc);
  return err;
}

bool block_tc_supported(const Step& s) {
  if (s.kind != STEP_BLOCK || s.w_umma < 0) return false;
  if (s.stride != 1 && s.stride != 2) return false;
  const int C = s.in.C, N = s.out.C;
  if (C % 8 != 0 || N % 4 != 0 || C < 16 || (kThreadsSmall % (C / 4)) != 0 || (kThreadsBig % (C / 4)) != 0) return false;
  if (s.Np > 128 || s.out.H < TH || s.out.W < TW) return false;
  if (s.stride == 1 && (s.pad_t != 1 || s.pad_l != 1 || s.in.H != s.out.H || s.in.W != s.out.W)) return false;
  if (s.stride == 2 && (s.pad_t != 0 || s.pad_l != 0 || s.in.H != 2 * s.out.H || s.in.W != 2 * s.out.W)) return false;
  if (s.in.offset != 0 || s.out.offset != 0 || s.in.batch_stride != (int64_t)s.in.H * s.in.W * C ||
      s.out.batch_stride != (int64_t)s.out.H * s.out.W * N)
    return false;
  if (s.skip.tensor >= 0 && (s.skip_c % 4 != 0 || s.skip.offset != 0)) return false;
  static const int max_n = getenv("FDL_BLOCK_TC_MAX_N") ? atoi(getenv("FDL_BLOCK_TC_MAX_N")) : 1 << 30;   // A/B timing against conv_tc
  if (N > max_n) return false;
  int stages, alias, total, cp, np;
  const int f16 = (tc_f16_enabled() && s.w_f16 >= 0) ? 1 : 0;
  return pick_smem(C, N, s.Np, s.stride, f16 ? s.wsplit16 : s.wsplit, f16, &stages, &alias, &total, &cp, &np);
}

cudaError_t launch_block_tc(const BlockTcLaunch& l, cudaStream_t stream) {
  if (!g_encode) return cudaErrorNotSupported;
  BlockTcArgs a = l.args;
  const int S = a.stride;
  int total = 0;
  a.f16 = (tc_f16_enabled() && a.w_f16 != nullptr) ? 1 : 0;
  if (!pick_smem(a.C, a.N, a.Np, S, a.f16 ? a.wsplit16 : a.wsplit, a.f16, &a.stages, &a.alias_out, &total, &a.tc_cp, &a.tc_np)) return cudaErrorInvalidConfiguration;
  a.pad = S == 1 ? 1 : 0;
  CUtensorMap tm_in, tm_out;
  if (!encode_nhwc(&tm_in, l.in, a.B, a.H * S, a.W * S, a.C, (long long)a.H * S * a.W * S * a.C, in_tile_h(S), in_tile_w(S), a.tc_cp))
    return cudaErrorInvalidValue;
  if (!encode_nhwc(&tm_out, l.out, a.B, a.H, a.W, a.N, (long long)a.H * a.W * a.N, TH, TW, a.tc_np)) return cudaErrorInvalidValue;
  CUtensorMap tm_skip = tm_in;                      // (a valid map when the residual does not travel by TMA)
  a.skip_tma = 0;
  static const bool skip_tma_on = getenv("FDL_TC_SKIP_TMA") ? atoi(getenv("FDL_TC_SKIP_TMA")) != 0 : true;
  if (skip_tma_on && a.skip_mode == 2 && !a.alias_out && a.skip != nullptr && a.skip_bstride == (long long)a.H * a.W * a.skip_c && a.skip_c % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a.skip) & 15) == 0 && encode_nhwc(&tm_skip, a.skip, a.B, a.H, a.W, a.skip_c, a.skip_bstride, TH, TW, a.tc_np))
    a.skip_tma = 1;
  a.tiles_x = (a.W + TW - 1) / TW;
  a.tiles_y = (a.H + TH - 1) / TH;
  a.tmem_cols = a.Np <= 32 ? 32 : (a.Np <= 64 ? 64 : 128);
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  int per_sm = (228 * 1024) / (total + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;                       // 168 registers x 192 threads: two resident CTAs per SM
  int grid = persist_sms() * per_sm;
  if (grid > ntiles) grid = ntiles;
  static const int big_env = getenv("FDL_TC_THREADS") ? atoi(getenv("FDL_TC_THREADS")) : kThreadsSmall;   // 384 measured no faster (r01v)
  const bool big = per_sm == 1 && big_env == kThreadsBig;
  cudaError_t e;
  if (S == 1 && big) e = launch_pdl(blaze_block_tc_kernel<1, kThreadsBig>, dim3(grid), dim3(kThreadsBig), (size_t)total, stream, tm_in, tm_out, tm_skip, a);
  else if (S == 1) e = launch_pdl(blaze_block_tc_kernel<1, kThreadsSmall>, dim3(grid), dim3(kThreadsSmall), (size_t)total, stream, tm_in, tm_out, tm_skip, a);
  else if (big) e = launch_pdl(blaze_block_tc_kernel<2, kThreadsBig>, dim3(grid), dim3(kThreadsBig), (size_t)total, stream, tm_in, tm_out, tm_skip, a);
  else e = launch_pdl(blaze_block_tc_kernel<2, kThreadsSmall>, dim3(grid), dim3(kThreadsSmall), (size_t)total, stream, tm_in, tm_out, tm_skip, a);
  count_launch();
  return e;
}

}  // namespace fdl
