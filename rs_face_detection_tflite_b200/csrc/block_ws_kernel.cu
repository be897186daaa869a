// block_ws_kernel.cu -- warp-specialised, software-pipelined BlazeBlock kernel for stride-1 blocks (sm_100a).
//
//     out = act( PW1x1( DW3x3(in) + b_dw ) + b_pw + skip )
//
// Same arithmetic as blaze_block_tc_kernel (mma_kernels.cu): depthwise 3x3 on the CUDA cores, pointwise 1x1 on the
// tensor cores (tcgen05.mma kind::tf32 with hi/lo operand splitting, accumulator in TMEM).
// What changes is the schedule.  One persistent CTA per SM runs two kinds of warps that work on DIFFERENT tiles at
// the same time, connected by mbarrier rings:
//
//   warps 4..   depthwise  G groups; group g takes the CTA's tiles g, g+G, ...: waits for the TMA'd 10 x 18 x C halo
//                          tile, 3x3 depthwise with packed FFMA2 (fma.rn.f32x2), tf32 hi/lo split, A operand written
//                          in the UMMA K-major core-matrix layout into the group's own A buffer.  The group's first
//                          thread then issues the tcgen05.mma chain for the tile into one of two TMEM accumulators
//                          (tcgen05.commit -> "accumulator full" and "A buffer free") and moves on to its next tile.
//   warps 0-3   epilogue   tcgen05.ld of the accumulator, + residual (prefetched from the resident input tile or from
//                          global memory), RELU / PRELU, 16-byte stores straight to global memory (NHWC).  When all
//                          four warps are done with a tile its input stage is refilled at once: thread 0 issues the
//                          TMA load of the tile NS places ahead.
//
// so the depthwise of tiles i+1..i+G overlaps the MMA and the epilogue of tile i, and NS tiles of TMA loads are in
// flight.  The pointwise bias rides on the tensor cores as one more K step (A rows (1,1,0,..), B rows (b_hi,b_lo,0,..)).
// Per tile the kernel reads the input tile once and writes the output tile once.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "mma_kernels.cuh"
#include "pdl.h"
#include "plan.h"
#include "sm100_ptx.cuh"

namespace fdl {

void count_launch();
bool encode_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, long long bstride, int box_h, int box_w, int box_c = 0);

// FDL_WS_TRACE (variant build only, `python -m rs_face_detection_tflite_b200.build --variant trace`): globaltimer stamps of the first CTAs'
// first tiles -- [cta][tile][event]: 0 load issued, 1 depthwise sees the tile, 2 depthwise thread 0 done, 3 all depthwise threads done,
// 4 MMAs issued, 5 epilogue sees the accumulator, 6 epilogue done with the tile, 7 the tile's TMA store issued.
#ifdef FDL_WS_TRACE
// (-DFDL_TRACE_C=32 -DFDL_TRACE_H=32: only launches of that block shape write the trace; default: every launch, the last one stays)
#if defined(FDL_TRACE_C) && defined(FDL_TRACE_H)
#define FDL_TRACE_MATCH (a.C == FDL_TRACE_C && a.H == FDL_TRACE_H)
#else
#define FDL_TRACE_MATCH true
#endif
constexpr int kTraceCtas = 4, kTraceTiles = 48;
__device__ unsigned long long g_ws_trace[kTraceCtas][kTraceTiles][8];
#define WS_T(it_, ev_)                                                                                              \
  do {                                                                                                              \
    if (blockIdx.x < kTraceCtas && (it_) >= 0 && (it_) < kTraceTiles && FDL_TRACE_MATCH) {                         \
      unsigned long long t_;                                                                                        \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                                        \
      g_ws_trace[blockIdx.x][(it_)][(ev_)] = t_;                                                                   \
    }                                                                                                               \
  } while (0)
cudaError_t ws_trace_read(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_ws_trace, sizeof(unsigned long long) * (size_t)(n < kTraceCtas * kTraceTiles * 8 ? n : kTraceCtas * kTraceTiles * 8));
}
#else
#define WS_T(it_, ev_) do { } while (0)
cudaError_t ws_trace_read(unsigned long long*, int) { return cudaErrorNotSupported; }
#endif

namespace {

typedef unsigned long long ull;

// Depthwise work assignment for Q = 6 channel quads (C = 24): item j of a half tile -> (quad | x << 4).  With the plain
// "quad fastest, then x" order a quarter warp's A-plane stores collide two-fold; this order (found by exhaustive search)
// keeps BOTH the 16-byte input loads ((6x + q) mod 8) and the plane stores ((q + x) mod 8) of every 8 consecutive lanes
// in 8 different bank groups.
__constant__ unsigned char kMapQ6[96] = {
    0, 16, 32, 48, 49, 65, 81, 97, 64, 80, 96, 112, 1, 17, 33, 113, 128, 144, 160, 176, 177, 193, 209, 225, 192, 208, 224, 240, 129, 145, 161, 241,
    2, 18, 34, 50, 51, 67, 83, 99, 66, 82, 98, 114, 3, 19, 35, 115, 130, 146, 162, 178, 179, 195, 211, 227, 194, 210, 226, 242, 131, 147, 163, 243,
    4, 20, 36, 52, 53, 69, 85, 101, 68, 84, 100, 116, 5, 21, 37, 117, 132, 148, 164, 180, 181, 197, 213, 229, 196, 212, 228, 244, 133, 149, 165, 245};

constexpr int TH = 8, TW = 16;                 // output tile: 128 pixels == UMMA M
constexpr int ITH = TH + 2, ITW = TW + 2;      // input halo tile
constexpr int kPlaneData = TH * TW * 16;      // one channel-quad plane of A: 128 pixels x 16 B
// Plane stride (== LBO) = kPlaneData + 16 B of bank skew between consecutive channel-quad planes.
__host__ __device__ inline int plane_bytes(int) { return kPlaneData + 16; }
constexpr int kEpiThreads = 128;               // 4 epilogue warps: one per TMEM lane quarter
constexpr int kMaxThreads = 512;              // epilogue + depthwise threads (the MMA warp comes on top)
constexpr int kMmaThreads = 32;
constexpr int kMaxStages = 6, kMaxGroups = 3, kMaxAcc = 4;
constexpr int kMaxSmemWs = 227 * 1024;

struct WsCfg { int G, ipt, ndwg, NS, OB, threads, total, ctas, in_pad, f16, teams; };
struct WsLayout { int alpha, w, wb, ones, in0, in_stage, a0, a_buf, out0, out_stage, total; };

__host__ __device__ inline int align_up_w(int v, int a) { return (v + a - 1) / a * a; }

// f16: the A operand is one plane set of (f16 hi, f16 lo) pairs (half the bytes), `wsplit` counts the f16 weight copies, and
// the bias is added in the epilogue (no bias K step: no `wb` / `ones` regions).
__host__ __device__ inline WsLayout ws_layout(int C, int N, int Np, int wsplit, int NS, int G, int OB, int in_pad, int f16) {
  WsLayout L;
  int off = 384;                                // barriers, tmem slot, descriptors
  L.alpha = off; off += Np * 4;
  off = align_up_w(off, 128);
  L.w = off; off += wsplit * (C / 4) * Np * 16;
  L.wb = off; if (!f16) off += 2 * Np * 16;     // bias as one more K step of B: [2 quads][Np][4] = (b_hi, b_lo, 0, 0), 0
  off = align_up_w(off, 128);
  L.ones = off; if (!f16) off += 2 * plane_bytes(C);   // the matching A planes: (1, 1, 0, 0) for every pixel, then zeros
  off = align_up_w(off, 128);
  L.in_stage = align_up_w(ITH * ITW * (in_pad ? ((C / 4) | 1) * 4 : C) * 4, 128);   // in_pad: pixel stride = odd number of 16-byte quads
  L.in0 = off; off += NS * L.in_stage;
  L.a_buf = align_up_w((f16 ? 1 : 2) * (C / 4) * plane_bytes(C), 128);   // tf32: hi planes then lo planes; f16: one set of (hi, lo) planes
  L.a0 = off; off += G * L.a_buf;
  L.out_stage = align_up_w(TH * TW * ((N / 4) | 1) * 16, 128);   // raw accumulator tile, pixel stride = odd number of quads
  L.out0 = off; off += OB * L.out_stage;
  L.total = align_up_w(off, 128);
  return L;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ ull fma2(ull a, ull b, ull c) {
  ull d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
// (c0, c1) -> f16x2 with c0 in the low half (the lower address), round to nearest even, saturating instead of overflowing
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

template <int kMaxT, int kMinB, bool kF16, bool kNarrow = false, bool kMmaWarp = false, int kTeams = 1>
__global__ void __launch_bounds__(kMaxT, kMinB) block_ws_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                                  const __grid_constant__ CUtensorMap tm_skip, const BlockTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kEpi = kTeams * kEpiThreads;    // epilogue threads: one or two teams of four warps (a warp per TMEM lane quarter)
  const int C = a.C, N = a.N, Np = a.Np, Q = C >> 2;
  const int NS = a.stages, G = a.groups, ndwg = a.dw_threads;
  const int kPlaneBytes = plane_bytes(C);
  constexpr bool f16 = kF16;                    // compile-time: the two operand formats must not share registers / code
  const int wcopies = f16 ? a.wsplit16 : a.wsplit;
  const WsLayout L = ws_layout(C, N, Np, wcopies, NS, G, a.out_bufs, a.in_pad, f16 ? 1 : 0);
  // TMEM accumulators: tile it -> buffer it % T.  Two epilogue teams: four, so that the depthwise / MMA side runs two tiles ahead of
  // each team instead of stalling on the drain of the tile two places back (every phase is tracked per tile: any issuer order works).
  const int T = kTeams == 2 ? kMaxAcc : (G > 2 ? G : 2);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem);             // [kMaxStages]  TMA tile landed
  uint64_t* a_full = in_full + kMaxStages;                           // [kMaxGroups]  A operand of the group written
  uint64_t* a_empty = a_full + kMaxGroups;                           // [kMaxGroups]  MMAs done reading the group's A buffer
  uint64_t* acc_full = a_empty + kMaxGroups;                         // [kMaxAcc]  accumulator complete
  uint64_t* acc_empty = acc_full + kMaxAcc;                          // [kMaxAcc]  accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kMaxAcc);
  uint64_t* skip_full = reinterpret_cast<uint64_t*>(smem + 168);      // [2] a team's residual tile landed in its staging buffer (two teams)
  uint64_t* s_desc = reinterpret_cast<uint64_t*>(smem + 192);         // [6] UMMA descriptors + [2] K-step increments, built once (group 0's)
  int* s_lstep = reinterpret_cast<int*>(smem + 256);                  // [3] tile-coordinate step of consecutive loads (image, row, column)
  float* s_w = reinterpret_cast<float*>(smem + L.w);

  const int tiles_per_img = a.tiles_x * a.tiles_y;

  const int CP = a.in_pad ? (((C >> 2) | 1) << 2) : C;   // pixel stride (floats) of the input tile in shared memory
  const uint32_t in_bytes = (uint32_t)(ITH * ITW * CP * 4);
  // Loads are issued by ONE thread (tid 0) strictly in tile order, so the tile coordinates advance by a fixed (image, row, column)
  // step per load: no divisions on the issuing thread's path (it sits in the epilogue's per-tile chain).
  int l_b = 0, l_ty = 0, l_tx = 0, l_s = 0;     // (live in tid 0 only)
  auto issue_load = [&](int it) {               // the CTA's it-th tile -> stage it % NS (called with it = 0, 1, 2, ... in order)
    ptx::mbar_arrive_expect_tx(&in_full[l_s], in_bytes);
    ptx::tma_load_4d(smem + L.in0 + l_s * L.in_stage, &tm_in, &in_full[l_s], 0, l_tx * TW - 1, l_ty * TH - 1, l_b);
    WS_T(it, 0);
    if (++l_s == NS) l_s = 0;
    l_tx += s_lstep[2]; l_ty += s_lstep[1]; l_b += s_lstep[0];
    if (l_tx >= a.tiles_x) { l_tx -= a.tiles_x; ++l_ty; }
    if (l_ty >= a.tiles_y) { l_ty -= a.tiles_y; ++l_b; }
  };

  // the same for any tile, by any thread (two epilogue teams: the refills are not issued in tile order by one thread)
  auto issue_load_at = [&](int it) {          // (divisions: one thread, once per tile)
    const int tile = (int)blockIdx.x + it * (int)gridDim.x, st = it % NS;
    const int lb = tile / tiles_per_img, r = tile - lb * tiles_per_img, lty = r / a.tiles_x, ltx = r - lty * a.tiles_x;
    ptx::mbar_arrive_expect_tx(&in_full[st], in_bytes);
    ptx::tma_load_4d(smem + L.in0 + st * L.in_stage, &tm_in, &in_full[st], 0, ltx * TW - 1, lty * TH - 1, lb);
    WS_T(it, 0);
  };

  // ---- one-time setup: nothing here depends on the previous launch (PDL, see pdl.h) ----
  if (tid == 0) {
    ptx::prefetch_tmap(&tm_in);
    ptx::prefetch_tmap(&tm_out);
    for (int s = 0; s < NS; ++s) ptx::mbar_init(&in_full[s], 1);
    for (int g = 0; g < G; ++g) { ptx::mbar_init(&a_full[g], (uint32_t)ndwg); ptx::mbar_init(&a_empty[g], 1); }
    for (int t = 0; t < T; ++t) { ptx::mbar_init(&acc_full[t], 1); ptx::mbar_init(&acc_empty[t], 4); }
    ptx::mbar_init(&skip_full[0], 1); ptx::mbar_init(&skip_full[1], 1);
    ptx::fence_mbar_init();
    const int t0 = (int)blockIdx.x, r0 = t0 % tiles_per_img;
    l_b = t0 / tiles_per_img; l_ty = r0 / a.tiles_x; l_tx = r0 - l_ty * a.tiles_x;
    const int sb = (int)gridDim.x / tiles_per_img, sr = (int)gridDim.x - sb * tiles_per_img;
    s_lstep[0] = sb; s_lstep[1] = sr / a.tiles_x; s_lstep[2] = sr - (sr / a.tiles_x) * a.tiles_x;
    const uint32_t w_a = ptx::smem_u32(smem + L.w), lbo_w = (uint32_t)Np * 16u;
    s_desc[0] = ptx::umma_desc_kmajor(ptx::smem_u32(smem + L.a0), kPlaneBytes, 128);
    s_desc[1] = ptx::umma_desc_kmajor(ptx::smem_u32(smem + L.a0) + (uint32_t)(Q * kPlaneBytes), kPlaneBytes, 128);
    s_desc[2] = ptx::umma_desc_kmajor(w_a, lbo_w, 128);
    s_desc[3] = ptx::umma_desc_kmajor(w_a + (uint32_t)(Q * Np * 16), lbo_w, 128);
    s_desc[4] = ptx::umma_desc_kmajor(ptx::smem_u32(smem + L.ones), kPlaneBytes, 128);
    s_desc[5] = ptx::umma_desc_kmajor(ptx::smem_u32(smem + L.wb), lbo_w, 128);
    s_desc[6] = (uint64_t)((2 * kPlaneBytes) >> 4);
    s_desc[7] = (uint64_t)((2u * lbo_w) >> 4);
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  if (!f16) {
    for (int i = tid; i < Np; i += blockDim.x) {
      const float bv = i < N ? a.bias[i] : 0.f;
      const float bh = __uint_as_float(__float_as_uint(bv) & 0xffffe000u);
      reinterpret_cast<float4*>(smem + L.wb)[i] = make_float4(bh, bv - bh, 0.f, 0.f);
      reinterpret_cast<float4*>(smem + L.wb)[Np + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < TH * TW; i += blockDim.x) {
      *reinterpret_cast<float4*>(smem + L.ones + i * 16) = make_float4(1.f, 1.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(smem + L.ones + kPlaneBytes + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  {  // pointwise weights: already in UMMA core-matrix order in global memory -> straight copy
    const int n4 = wcopies * Q * Np;
    const float4* src = reinterpret_cast<const float4*>(f16 ? a.w_f16 : a.w_umma);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (provably warp-uniform for the MMA warp)
  const int acc_cols = a.acc_cols;              // columns per accumulator buffer
  pdl_launch_dependents();
  pdl_wait();                                   // the previous launch's activations (and *n_active) are visible from here on
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int ntiles = nb * tiles_per_img;
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (tid == 0)
    for (int it = 0; it < NS && it < my_tiles; ++it) issue_load(it);

  if (my_tiles == 0) {
    // nothing to do (fewer active items than CTAs)
  } else if (warp < 4 * kTeams) {
    // ================= epilogue: TMEM -> (+skip, act) -> staging tile -> TMA store; refills the input ring =================
    // thread == TMEM lane == pixel.  Both the input tile (TMA load with a box wider than the tensor: the tail of each
    // pixel is zero-filled) and the output staging tile (TMA store with the same trick: the tail is clipped) have a
    // pixel stride of an ODD number of 16-byte quads, so the per-pixel 16-byte accesses are bank-conflict free.
    const int p = tid & 127;                    // TMEM lane (warp w may access lanes 32 (w % 4) .. + 31)
    const int py = p / TW, px = p - py * TW;
    const int NPf = ((N >> 2) | 1) << 2;        // staging pixel stride (floats)
    int pb = 0, pty = 0, ptx_ = 0;              // coordinates of the previous tile (its TMA store is issued one tile late)
    // tile coordinates advance by a fixed (image, row, column) step per iteration: no divisions inside the loop
    int b, ty, tx;
    {
      const int t0 = (int)blockIdx.x, r0 = t0 % tiles_per_img;
      b = t0 / tiles_per_img; ty = r0 / a.tiles_x; tx = r0 - ty * a.tiles_x;
    }
    const int step_b = (int)gridDim.x / tiles_per_img, step_r = (int)gridDim.x - step_b * tiles_per_img;
    const int step_y = step_r / a.tiles_x, step_x = step_r - step_y * a.tiles_x;
    // ---- narrow blocks (N <= 32, residual from the resident input tile, two staging buffers): the epilogue sets the CTA's pace (measured
    // with FDL_WS_TRACE: ~1.9 us per tile, most of it index arithmetic, runtime-bounded loops and one TMEM round trip per 16
    // columns), so this path keeps its indices in counters, its residual in six registers and both accumulator halves in flight.
    // ---- two epilogue teams (wide blocks, one CTA per SM): on the 32 x 32 x 48 detector stage the timeline (FDL_WS_TRACE) has the
    // depthwise group done with a tile in 0.65 us and ONE team of four warps busy 2.4 us per tile (1.5 us of epilogue arithmetic for
    // 128 pixels x 48 channels, the store hand-over, the refill): the epilogue sets the pace.  Two teams take alternate tiles; each
    // owns a staging buffer, issues its own TMA stores and refills the input stage of the tile it has just finished.
    if constexpr (kTeams == 2) {
      const int team = warp >> 2, bar_id = 1 + team;
      const bool leader = p == 0;
      const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      float* s_o = reinterpret_cast<float*>(smem + L.out0 + team * L.out_stage) + p * NPf;
      const bool skip_tma = a.skip_mode == 2 && a.skip_tma != 0;
      // No staging tile at all (out_bufs == 0: wide outputs whose two 67 KB staging tiles would not fit): a thread owns a pixel, i.e.
      // N contiguous floats of the NHWC output, and stores them straight to global memory, whole 128-byte lines over the column loop.
      const bool direct = a.out_bufs == 0;
      for (int it = 0; it < my_tiles; ++it) {
        if (it > 0) {
          tx += step_x; ty += step_y; b += step_b;
          if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
          if (ty >= a.tiles_y) { ty -= a.tiles_y; ++b; }
        }
        if ((it & 1) != team) continue;
        const int s = it % NS, t = it % T;
        const int oy = ty * TH + py, ox = tx * TW + px;
        const bool inside = oy < a.H && ox < a.W;
        const float* skip_smem = reinterpret_cast<const float*>(smem + L.in0 + s * L.in_stage) + ((py + 1) * ITW + (px + 1)) * CP;
        const float* skip_g = (a.skip_mode == 2 && inside) ? a.skip + (long long)b * a.skip_bstride + ((long long)oy * a.W + ox) * a.skip_c : nullptr;
        // MAX_POOL 2x2 of a map twice as large (skip_mode 3): the window's four pixels, per-thread loads
        const float* skip_p = (a.skip_mode == 3 && inside) ? a.skip + (long long)b * a.skip_bstride + ((long long)(2 * oy) * (2 * a.W) + 2 * ox) * a.skip_c : nullptr;
        const long long prs = (long long)2 * a.W * a.skip_c;
        float4 res[8], resn[8];
        auto load_res = [&](int c0, float4 (&d)[8]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = c0 + 4 * j;
            d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (skip_tma) { if (n < N) d[j] = ld4(s_o + n); }
            else if (n < a.skip_c) {
              if (a.skip_mode == 1) d[j] = ld4(skip_smem + n);
              else if (skip_g) d[j] = __ldg(reinterpret_cast<const float4*>(skip_g + n));
              else if (skip_p)
                d[j] = max4(max4(__ldg(reinterpret_cast<const float4*>(skip_p + n)), __ldg(reinterpret_cast<const float4*>(skip_p + a.skip_c + n))),
                            max4(__ldg(reinterpret_cast<const float4*>(skip_p + prs + n)), __ldg(reinterpret_cast<const float4*>(skip_p + prs + a.skip_c + n))));
            }
          }
        };
        if (skip_tma) {
          // the residual tile travels by TMA straight into the team's staging buffer (same pixel stride as the output tile; channels
          // beyond skip_c are zero-filled: the channel PAD) and the epilogue updates it in place -- no per-thread global loads
          if (leader) {
            ptx::tma_store_wait_read0();                // this team's previous store has read its staging buffer
            ptx::mbar_arrive_expect_tx(&skip_full[team], (uint32_t)(TH * TW * NPf * 4));
            ptx::tma_load_4d(smem + L.out0 + team * L.out_stage, &tm_skip, &skip_full[team], 0, tx * TW, ty * TH, b);
          }
          ptx::mbar_wait(&skip_full[team], (uint32_t)((it >> 1) & 1));
        }
        if (a.skip_mode == 1) ptx::mbar_wait(&in_full[s], (uint32_t)((it / NS) & 1));
        load_res(0, res);
        ptx::mbar_wait(&acc_full[t], (uint32_t)((it / T) & 1));
        ptx::tc_fence_after_sync();
        if (tid == 0) WS_T(it, 5);
        if (!skip_tma && !direct) {
          if (leader) ptx::tma_store_wait_read0();      // this team's previous store has read its staging buffer
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        float* out_g = (direct && inside) ? a.out_direct + (((long long)b * a.H + oy) * a.W + ox) * N : nullptr;
        const uint32_t taddr = taddr0 + (uint32_t)(t * acc_cols);
        for (int c0 = 0; c0 < Np; c0 += 32) {
          if (c0 + 32 < Np) load_res(c0 + 32, resn);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = c0 + 16 * h;
            if (ch >= Np) break;
            uint32_t r[16];
            ptx::tmem_ld16_issue(taddr + (uint32_t)ch, r);
            ptx::tmem_ld_wait16(r);
            if (ch + 16 >= Np) {                    // accumulator drained: hand it back to the issuing thread
              ptx::tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[t]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = ch + 4 * j;
              if (n >= N) break;
              float4 rs = res[4 * h + j];
              if (f16) { rs.x += a.bias_c[n]; rs.y += a.bias_c[n + 1]; rs.z += a.bias_c[n + 2]; rs.w += a.bias_c[n + 3]; }
              float4 o = make_float4(__uint_as_float(r[4 * j]) + rs.x, __uint_as_float(r[4 * j + 1]) + rs.y, __uint_as_float(r[4 * j + 2]) + rs.z,
                                     __uint_as_float(r[4 * j + 3]) + rs.w);
              if (a.act == ACT_RELU) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              } else if (a.act == ACT_PRELU) {
                o.x = o.x >= 0.f ? o.x : o.x * a.alpha_c[n]; o.y = o.y >= 0.f ? o.y : o.y * a.alpha_c[n + 1];
                o.z = o.z >= 0.f ? o.z : o.z * a.alpha_c[n + 2]; o.w = o.w >= 0.f ? o.w : o.w * a.alpha_c[n + 3];
              }
              if (direct) { if (out_g) *reinterpret_cast<float4*>(out_g + n) = o; }
              else *reinterpret_cast<float4*>(s_o + n) = o;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = resn[j];
        }
        if (tid == 0) WS_T(it, 6);
        if (!direct) ptx::fence_proxy_async_smem();
        if (!direct || a.skip_mode == 1) asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (leader) {
          if (!direct) {
            ptx::tma_store_4d(&tm_out, smem + L.out0 + team * L.out_stage, 0, tx * TW, ty * TH, b);
            ptx::tma_store_commit();
          }
          // residual from the resident input tile: the team is past its reads of this tile's stage (and the depthwise finished with it
          // before the MMAs): refill.  (Otherwise the depthwise group refills as soon as IT is done with the stage, see there.)
          if (a.skip_mode == 1 && it + NS < my_tiles) issue_load_at(it + NS);
        }
      }
      if (leader && !direct) ptx::tma_store_wait_all0();
    } else
    if constexpr (kNarrow) {
      int s = 0, t = 0, ph_in = 0, ph_acc = 0;
      const int sq = a.skip_c >> 2;
      const bool relu = a.act == ACT_RELU, prelu = a.act == ACT_PRELU;
      const float* skip0 = reinterpret_cast<const float*>(smem + L.in0) + ((py + 1) * ITW + (px + 1)) * CP;
      float* s_o0 = reinterpret_cast<float*>(smem + L.out0) + p * NPf;
      const int in_stage_f = L.in_stage >> 2, out_stage_f = L.out_stage >> 2;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int it = 0; it < my_tiles; ++it) {
        const int ob = it & 1;
        if (it > 0) {
          tx += step_x; ty += step_y; b += step_b;
          if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
          if (ty >= a.tiles_y) { ty -= a.tiles_y; ++b; }
        }
        ptx::mbar_wait(&in_full[s], (uint32_t)ph_in);
        const float* sk = skip0 + s * in_stage_f;
        float4 res[4];                                        // residual of the first 16 channels; the second half follows below
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = j < sq ? ld4(sk + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        ptx::mbar_wait(&acc_full[t], (uint32_t)ph_acc);
        ptx::tc_fence_after_sync();
        if (tid == 0) WS_T(it, 5);
        uint32_t r0[16], r1[16];
        const uint32_t taddr = taddr0 + (uint32_t)(t * acc_cols);
        ptx::tmem_ld16_issue(taddr, r0);
        if (Np > 16) ptx::tmem_ld16_issue(taddr + 16u, r1);
        ptx::fence_proxy_async_smem();
        if (tid == 0) ptx::tma_store_wait_read0();            // the store of tile it-2 has finished reading buffer `ob`
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          if (it > 0) {
            ptx::tma_store_4d(&tm_out, smem + L.out0 + (ob ^ 1) * L.out_stage, 0, ptx_ * TW, pty * TH, pb);
            ptx::tma_store_commit();
            WS_T(it - 1, 7);
          }
          // every epilogue warp is past its residual reads of the previous tile's stage: refill it.  (Refilling the CURRENT tile's
          // stage one tile earlier was measured again in round 2, with this epilogue and the f16 operand: 217 vs 192 us.  More loads
          // in flight make the launch slower, as in round 1: the memory system prefers the shallower queue.)
          if (it > 0 && it - 1 + NS < my_tiles) issue_load(it - 1 + NS);
        }
        pb = b; pty = ty; ptx_ = tx;
        ptx::tmem_ld_wait16(r0);
        if (Np > 16) ptx::tmem_ld_wait16(r1);
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[t]);
        float* so = s_o0 + ob * out_stage_f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = 16 * h + 4 * j;
            if (n < N) {
              const uint32_t* r = h ? r1 : r0;
              float4 rs = res[j];
              if (f16) { rs.x += a.bias_c[n]; rs.y += a.bias_c[n + 1]; rs.z += a.bias_c[n + 2]; rs.w += a.bias_c[n + 3]; }
              float4 o = make_float4(__uint_as_float(r[4 * j]) + rs.x, __uint_as_float(r[4 * j + 1]) + rs.y, __uint_as_float(r[4 * j + 2]) + rs.z,
                                     __uint_as_float(r[4 * j + 3]) + rs.w);
              if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              else if (prelu) {                       // slopes come from the kernel parameters (constant bank)
                o.x = o.x >= 0.f ? o.x : o.x * a.alpha_c[n]; o.y = o.y >= 0.f ? o.y : o.y * a.alpha_c[n + 1];
                o.z = o.z >= 0.f ? o.z : o.z * a.alpha_c[n + 2]; o.w = o.w >= 0.f ? o.w : o.w * a.alpha_c[n + 3];
              }
              *reinterpret_cast<float4*>(so + n) = o;
            }
          }
          if (h == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = 4 + j < sq ? ld4(sk + 16 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (tid == 0) WS_T(it, 6);
        if (++s == NS) { s = 0; ph_in ^= 1; }
        if (++t == T) { t = 0; ph_acc ^= 1; }
      }
    } else {
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NS, t = it % T, ob = a.out_bufs == 2 ? (it & 1) : 0;
      if (it > 0) {
        tx += step_x; ty += step_y; b += step_b;
        if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
        if (ty >= a.tiles_y) { ty -= a.tiles_y; ++b; }
      }
      const int oy = ty * TH + py, ox = tx * TW + px;
      const bool inside = oy < a.H && ox < a.W;
      const float* skip_smem = reinterpret_cast<const float*>(smem + L.in0 + s * L.in_stage) + ((py + 1) * ITW + (px + 1)) * CP;
      const float* skip_g = (a.skip_mode == 2 && inside) ? a.skip + (long long)b * a.skip_bstride + ((long long)oy * a.W + ox) * a.skip_c : nullptr;
      float* s_o = reinterpret_cast<float*>(smem + L.out0 + ob * L.out_stage) + p * NPf;
      // Residual of 32 channels at a time.  The first batch is requested BEFORE waiting for the accumulator: the
      // shared-memory pipe is kept saturated by the depthwise warps, so a load issued here takes hundreds of cycles.
      float4 res[8], resn[8];
      auto load_res = [&](int c0, float4 (&d)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = c0 + 4 * j;
          d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < a.skip_c) {
            if (a.skip_mode == 1) d[j] = ld4(skip_smem + n);
            else if (skip_g) d[j] = __ldg(reinterpret_cast<const float4*>(skip_g + n));
          }
        }
      };
      if (a.skip_mode == 1) ptx::mbar_wait(&in_full[s], (uint32_t)((it / NS) & 1));
      load_res(0, res);
      ptx::mbar_wait(&acc_full[t], (uint32_t)((it / T) & 1));
      ptx::tc_fence_after_sync();
      if (tid == 0) WS_T(it, 5);
      // ---- deferred TMA store of the PREVIOUS tile: its staging writes have had a whole tile to drain ----
      // (one staging buffer: the store is issued here too, and this tile's staging writes wait below until it has read the buffer)
      {
        ptx::fence_proxy_async_smem();
        if (tid == 0 && a.out_bufs == 2) ptx::tma_store_wait_read0();   // the store of tile it-2 has finished reading buffer `ob`
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0 && it > 0) {
          ptx::tma_store_4d(&tm_out, smem + L.out0 + (a.out_bufs == 2 ? (ob ^ 1) : 0) * L.out_stage, 0, ptx_ * TW, pty * TH, pb);
          ptx::tma_store_commit();
          WS_T(it - 1, 7);
          // every epilogue warp is past its residual reads of the previous tile's stage: refill it
          if (it - 1 + NS < my_tiles) issue_load(it - 1 + NS);
        }
      }
      pb = b; pty = ty; ptx_ = tx;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * acc_cols);
      bool staging_free = a.out_bufs == 2;
      for (int c0 = 0; c0 < Np; c0 += 32) {
        if (c0 + 32 < Np) load_res(c0 + 32, resn);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = c0 + 16 * h;
          if (ch >= Np) break;
          uint32_t r[16];
          ptx::tmem_ld16_issue(taddr + (uint32_t)ch, r);
          ptx::tmem_ld_wait16(r);
          if (!staging_free) {                    // single staging buffer: the previous tile's store (issued above) must have read it
            if (tid == 0) ptx::tma_store_wait_read0();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            staging_free = true;
          }
          if (ch + 16 >= Np) {                    // accumulator drained: hand it back to the issuing thread
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[t]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = ch + 4 * j;
            if (n >= N) break;
            float4 rs = res[4 * h + j];
            if (f16) { rs.x += a.bias_c[n]; rs.y += a.bias_c[n + 1]; rs.z += a.bias_c[n + 2]; rs.w += a.bias_c[n + 3]; }   // tf32 mode: bias came as a K step
            float4 o = make_float4(__uint_as_float(r[4 * j]) + rs.x, __uint_as_float(r[4 * j + 1]) + rs.y, __uint_as_float(r[4 * j + 2]) + rs.z,
                                   __uint_as_float(r[4 * j + 3]) + rs.w);
            if (a.act == ACT_RELU) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            } else if (a.act == ACT_PRELU) {      // slopes come from the kernel parameters (constant bank), not shared memory
              o.x = o.x >= 0.f ? o.x : o.x * a.alpha_c[n]; o.y = o.y >= 0.f ? o.y : o.y * a.alpha_c[n + 1];
              o.z = o.z >= 0.f ? o.z : o.z * a.alpha_c[n + 2]; o.w = o.w >= 0.f ? o.w : o.w * a.alpha_c[n + 3];
            }
            *reinterpret_cast<float4*>(s_o + n) = o;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) res[j] = resn[j];
      }
      if (tid == 0) WS_T(it, 6);
    }
    }
    if constexpr (kTeams == 1) {
    // the last tile's store
    ptx::fence_proxy_async_smem();
    if (tid == 0) ptx::tma_store_wait_read0();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
      ptx::tma_store_4d(&tm_out, smem + L.out0 + (a.out_bufs == 2 ? ((my_tiles - 1) & 1) : 0) * L.out_stage, 0, ptx_ * TW, pty * TH, pb);
      ptx::tma_store_commit();
      ptx::tma_store_wait_all0();
    }
    }
  } else if (tid - kEpi < G * ndwg) {
    // ================= depthwise 3x3 -> A operand (hi / lo planes) =================
    const int dtid = tid - kEpi;
    const int g = dtid / ndwg, gt = dtid - g * ndwg;
    // Q == 6, one item per thread: with the plain 6-quad pixel stride the items are remapped for conflict-free stores (kMapQ6);
    // with the padded 7-quad stride (f16 mode: shared memory allows it) "x fastest" is conflict-free for loads AND stores.
    const bool xfast = Q == 6 && (ndwg == 192 || ndwg == 96) && a.in_pad;
    const bool map6 = Q == 6 && (ndwg == 192 || ndwg == 96) && !a.in_pad;
    const int m6 = map6 ? kMapQ6[gt % 96] : 0;
    const int q = map6 ? (m6 & 15) : (xfast ? (gt / TW) % Q : gt % Q);    // the channel quad is fixed per thread
    ull wd[9][2], bd[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2*>(a.w_dw + k * C) + q);
      wd[k][0] = w.x; wd[k][1] = w.y;
    }
    {
      const ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2*>(a.b_dw) + q);
      bd[0] = w.x; bd[1] = w.y;
    }
    const ull kMaskHi = 0xffffe000ffffe000ull;
    const float2 neg1f = make_float2(-1.f, -1.f);
    const ull kNeg1 = *reinterpret_cast<const ull*>(&neg1f);
    uint8_t* s_ahi = smem + L.a0 + g * L.a_buf;
    uint8_t* s_alo = s_ahi + Q * kPlaneBytes;   // tf32 mode only
    const int nitems = Q * 2 * TW;
    // MMA issue state (used by the group's first thread only)
    const uint32_t idesc = f16 ? ptx::umma_idesc_f16(128, Np) : ptx::umma_idesc_tf32(128, Np);
    // The MMA chain is issued by the group's first thread between two of its own depthwise items: every instruction it spends
    // there delays the whole group's next tile (measured: 0.75 us per tile with the descriptors rebuilt per instruction).  The
    // descriptors are built once; a K step only adds to their 14-bit address fields.
    // (kept in shared memory -- only the issuing thread needs them, and the depthwise loop is at the register cap -- as group 0's
    // A descriptors; group g's differ by its buffer offset in the address field)
    const uint64_t a_goff = (uint64_t)((uint32_t)(g * L.a_buf) >> 4);
    for (int it = g; it < my_tiles; it += G) {
      const int s = it % NS, kg = it / G, t = it % T;
      ptx::mbar_wait(&in_full[s], (uint32_t)((it / NS) & 1));
      if (gt == 0) WS_T(it, 1);
      const float* s_in = reinterpret_cast<const float*>(smem + L.in0 + s * L.in_stage);
      for (int item = gt, ii = 0; item < nitems; item += ndwg, ++ii) {
        int x, half;                            // (runtime divisions by Q only on the generic path)
        if (map6) { x = m6 >> 4; half = item / 96; }
        else if (xfast) { x = item % TW; half = item / 96; }
        else { const int xr = item / Q; x = xr % TW; half = xr / TW; }   // item % Q == q
        ull acc[4][2];
#pragma unroll
        for (int o = 0; o < 4; ++o) { acc[o][0] = bd[0]; acc[o][1] = bd[1]; }
        const float* base = s_in + ((half * 4) * ITW + x) * CP + 4 * q;
        ulonglong2 v[3][3];                     // ring of three input rows (x, x+1, x+2): loads run two rows ahead
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float* rp = base + r * ITW * CP;
          v[r][0] = *reinterpret_cast<const ulonglong2*>(rp);
          v[r][1] = *reinterpret_cast<const ulonglong2*>(rp + CP);
          v[r][2] = *reinterpret_cast<const ulonglong2*>(rp + 2 * CP);
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {           // input rows feeding 4 consecutive output rows
          if (r + 2 < 6) {
            const float* rp = base + (r + 2) * ITW * CP;
            v[(r + 2) % 3][0] = *reinterpret_cast<const ulonglong2*>(rp);
            v[(r + 2) % 3][1] = *reinterpret_cast<const ulonglong2*>(rp + CP);
            v[(r + 2) % 3][2] = *reinterpret_cast<const ulonglong2*>(rp + 2 * CP);
          }
          const ulonglong2 v0 = v[r % 3][0], v1 = v[r % 3][1], v2 = v[r % 3][2];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int ky = r - o;
            if (ky >= 0 && ky < 3) {
              acc[o][0] = fma2(v0.x, wd[ky * 3 + 0][0], acc[o][0]); acc[o][1] = fma2(v0.y, wd[ky * 3 + 0][1], acc[o][1]);
              acc[o][0] = fma2(v1.x, wd[ky * 3 + 1][0], acc[o][0]); acc[o][1] = fma2(v1.y, wd[ky * 3 + 1][1], acc[o][1]);
              acc[o][0] = fma2(v2.x, wd[ky * 3 + 2][0], acc[o][0]); acc[o][1] = fma2(v2.y, wd[ky * 3 + 2][1], acc[o][1]);
            }
          }
        }
        // the MMAs of this group's previous tile must have finished reading the A buffer
        if (ii == 0 && kg > 0) ptx::mbar_wait(&a_empty[g], (uint32_t)((kg - 1) & 1));
        if (f16) {
          // (f16 hi, f16 lo) of the four channels in ONE 16-byte core-matrix row: K' = (hi c0..c3, lo c0..c3)
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int p = (half * 4 + o) * TW + x;
            const float2 c01 = *reinterpret_cast<const float2*>(&acc[o][0]), c23 = *reinterpret_cast<const float2*>(&acc[o][1]);
            uint4 v;
            v.x = pack_f16x2(c01.x, c01.y); v.y = pack_f16x2(c23.x, c23.y);
            const float2 h01 = unpack_f16x2(v.x), h23 = unpack_f16x2(v.y);
            v.z = pack_f16x2(c01.x - h01.x, c01.y - h01.y); v.w = pack_f16x2(c23.x - h23.x, c23.y - h23.y);
            *reinterpret_cast<uint4*>(s_ahi + q * kPlaneBytes + p * 16) = v;
          }
        } else {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int p = (half * 4 + o) * TW + x;
            ulonglong2 hi, lo;
            hi.x = acc[o][0] & kMaskHi; hi.y = acc[o][1] & kMaskHi;
            lo.x = fma2(hi.x, kNeg1, acc[o][0]); lo.y = fma2(hi.y, kNeg1, acc[o][1]);
            *reinterpret_cast<ulonglong2*>(s_ahi + q * kPlaneBytes + p * 16) = hi;
            *reinterpret_cast<ulonglong2*>(s_alo + q * kPlaneBytes + p * 16) = lo;
          }
        }
      }
      ptx::fence_proxy_async_smem();            // generic-proxy writes of A -> visible to the tensor core (async proxy)
      mbar_arrive(&a_full[g]);
      if (gt == 0) WS_T(it, 2);
      if constexpr (!kMmaWarp) {
      if (gt < 32) {
        // ---- this tile's MMA chain: the group's first warp, converged; one elected lane issues (operands stay uniform: no per-
        // instruction ELECT / R2UR retry loop as in an `if (thread == 0)` region) ----
        ptx::mbar_wait(&a_full[g], (uint32_t)(kg & 1));
        if (gt == 0) WS_T(it, 3);
        if constexpr (kTeams == 2) {
          // nobody else reads this tile's input stage (residual from global memory, or none): refill it now, not after the epilogue --
          // on the iris 32 x 32 blocks the trace had the next load issued 5 us late and the depthwise waiting 3 us for it
          if (a.skip_mode != 1 && gt == 0 && it + NS < my_tiles) issue_load_at(it + NS);
          __syncwarp();
        }
        if (it >= T) ptx::mbar_wait(&acc_empty[t], (uint32_t)(((it / T) - 1) & 1));
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(t * acc_cols);
        const uint64_t d_ahi = s_desc[0] + a_goff, d_alo = s_desc[1] + a_goff, d_w = s_desc[2], d_w2 = s_desc[3], a_step = s_desc[6], b_step = s_desc[7];
        if (f16) {
          // kind::f16, K = 16 = two planes = two channel quads x (hi, lo); pass 0: (A_hi, A_lo) * (W, W), pass 1: (A_hi, A_lo) * (W_lo, 0)
          const int ksteps16 = Q >> 1;
          for (int pass = 0; pass < a.wsplit16; ++pass) {
            uint64_t ad = d_ahi, bdsc = pass ? d_w2 : d_w;
            for (int ks = 0; ks < ksteps16; ++ks, ad += a_step, bdsc += b_step) ptx::mma_f16_elect(d_tmem, ad, bdsc, idesc, (pass | ks) ? 1u : 0u);
          }
          ptx::mma_commit_elect(&acc_full[t]);
          ptx::mma_commit_elect(&a_empty[g]);
          if (gt == 0) WS_T(it, 4);
          continue;
        }
        // bias K step first (overwrites the accumulator), then the hi / lo passes accumulate
        ptx::mma_tf32_elect(d_tmem, s_desc[4], s_desc[5], idesc, 0u);
        const int ksteps = C >> 3;
        const int npass = a.wsplit == 2 ? 3 : 2;
        for (int pass = 0; pass < npass; ++pass) {
          // pass 0: A_lo * W_hi, pass 1: A_hi * W_hi, pass 2: A_hi * W_lo
          uint64_t ad = pass == 0 ? d_alo : d_ahi, bdsc = pass == 2 ? d_w2 : d_w;
          for (int ks = 0; ks < ksteps; ++ks, ad += a_step, bdsc += b_step) ptx::mma_tf32_elect(d_tmem, ad, bdsc, idesc, 1u);
        }
        ptx::mma_commit_elect(&acc_full[t]);
        ptx::mma_commit_elect(&a_empty[g]);
        if (gt == 0) WS_T(it, 4);
      }
      }
    }
  } else if constexpr (kMmaWarp) {
    // ================= the last warp: MMA issue.  The whole warp runs the loop converged and one elected lane issues, so the
    // operands stay on the uniform datapath (an `if (thread == 0)` region makes ptxas move every operand register -> uniform
    // register with an ELECT / R2UR retry loop per instruction, ~50 ns each); and no depthwise warp is held up by the issue.
    // Used where the extra warp costs no registers (C = 16: 9 warps per CTA): 119 -> 100 us on the 96 x 96 x 16 blocks.  For C = 24
    // (11 warps, 96 -> 80 registers) it measured slower (192 -> 200 us), so there the depthwise group's first thread issues. =================
    const uint32_t idesc = f16 ? ptx::umma_idesc_f16(128, Np) : ptx::umma_idesc_tf32(128, Np);
    const uint64_t a_step = s_desc[6], b_step = s_desc[7];
    for (int it = 0; it < my_tiles; ++it) {
      const int g = it % G, kg = it / G, t = it % T;
      ptx::mbar_wait(&a_full[g], (uint32_t)(kg & 1));
      if (lane == 0) WS_T(it, 3);
      if (it >= T) ptx::mbar_wait(&acc_empty[t], (uint32_t)(((it / T) - 1) & 1));
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(t * acc_cols);
      const uint64_t a_goff = (uint64_t)((uint32_t)(g * L.a_buf) >> 4);
      const uint64_t d_ahi = s_desc[0] + a_goff, d_alo = s_desc[1] + a_goff, d_w = s_desc[2], d_w2 = s_desc[3];
      if (f16) {
        // kind::f16, K = 16 = two planes = two channel quads x (hi, lo); pass 0: (A_hi, A_lo) * (W, W), pass 1: (A_hi, A_lo) * (W_lo, 0)
        const int ksteps16 = Q >> 1;
        for (int pass = 0; pass < a.wsplit16; ++pass) {
          uint64_t ad = d_ahi, bdsc = pass ? d_w2 : d_w;
          for (int ks = 0; ks < ksteps16; ++ks, ad += a_step, bdsc += b_step) ptx::mma_f16_elect(d_tmem, ad, bdsc, idesc, (pass | ks) ? 1u : 0u);
        }
      } else {
        // bias K step first (overwrites the accumulator), then the hi / lo passes accumulate
        ptx::mma_tf32_elect(d_tmem, s_desc[4], s_desc[5], idesc, 0u);
        const int ksteps = C >> 3;
        const int npass = a.wsplit == 2 ? 3 : 2;
        for (int pass = 0; pass < npass; ++pass) {
          // pass 0: A_lo * W_hi, pass 1: A_hi * W_hi, pass 2: A_hi * W_lo
          uint64_t ad = pass == 0 ? d_alo : d_ahi, bdsc = pass == 2 ? d_w2 : d_w;
          for (int ks = 0; ks < ksteps; ++ks, ad += a_step, bdsc += b_step) ptx::mma_tf32_elect(d_tmem, ad, bdsc, idesc, 1u);
        }
      }
      ptx::mma_commit_elect(&acc_full[t]);
      ptx::mma_commit_elect(&a_empty[g]);
      if (lane == 0) WS_T(it, 4);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// Depthwise thread organisation per channel-quad count: G groups of ndwg threads, ipt items per thread and tile, and
// how many CTAs share an SM.  Small channel counts (little work per tile, latency-dominated) run TWO smaller CTAs per SM
// so that two epilogues and two depthwise groups are in flight; larger ones run one CTA with up to 12 depthwise warps.
// f16-split A operand (FDL_WS_F16=1; off by default): see BlockTcArgs::f16.  It removes 27 % of the kernel's shared-memory
// wavefronts and all its bank conflicts, and measures no faster (229.9 vs 227.6 us on the 128x128x24 stage, 47 vs 45 us at 32x32x48):
// the kernel is not bound by shared-memory bandwidth.  The serial kernel does use it (room for a second stage).
bool ws_f16_enabled() {
  static const bool on = [] { const char* e = getenv("FDL_WS_F16"); return e ? atoi(e) != 0 : true; }();
  return on;
}

bool pick_cfg(int C, int N, int Np, int wsplit, int f16, WsCfg* cfg) {
  const int Q = C / 4;
  int G = 1, ipt = 1, ctas = 1;
  // FDL_WS_CTAS=3 (f16 operand only): three smaller CTAs per SM for the narrow blocks -- 96 depthwise threads (two items each), two input
  // stages, one staging buffer: six tiles in flight per SM instead of four.
  static const int ctas_env = getenv("FDL_WS_CTAS") ? atoi(getenv("FDL_WS_CTAS")) : 2;
  const bool three = ctas_env == 3 && f16 && (Q == 4 || Q == 6);
  // Wide blocks (one CTA per SM): two epilogue teams, paid for with depthwise threads (see the kernel).  FDL_WS_TEAMS=1: one team.
  static const int teams_env = getenv("FDL_WS_TEAMS") ? atoi(getenv("FDL_WS_TEAMS")) : 2;
  const int teams = (Q == 4 || Q == 6 || teams_env != 2) ? 1 : 2;
  const int epi = teams * kEpiThreads;
  switch (Q) {
    case 4: G = 1; ipt = three ? 2 : 1; ctas = three ? 3 : 2; break;
    case 6: G = 1; ipt = three ? 2 : 1; ctas = three ? 3 : 2; break;
    case 8: G = 3; ipt = 2; break;
    default:
      ipt = 1;
      while ((32 * Q) / ipt > kMaxThreads - epi || (32 * Q) % ipt != 0 || ((32 * Q) / ipt) % 32 != 0 || ((32 * Q) / ipt) % Q != 0) {
        if (++ipt > 8) return false;
      }
      break;
  }
  // Padding the pixel stride to an odd number of quads makes the epilogue's residual reads conflict-free.  The depthwise loads
  // stay conflict-free with it when Q % 8 == 0 (quad-fastest items) and, for Q == 6, with "x fastest" items -- which only fits
  // next to two CTAs per SM in f16 mode (smaller A operand, no bias planes).
  static const int pad6 = getenv("FDL_WS_PAD6") ? atoi(getenv("FDL_WS_PAD6")) : 1;   // A/B: padded stride + 3 stages vs plain stride + 4 stages
  const int in_pad = (Q % 8 == 0 || (f16 && Q == 6 && pad6)) ? 1 : 0;
  static const int ns_cap = getenv("FDL_WS_NS") ? atoi(getenv("FDL_WS_NS")) : 0;
  const int budget = ctas == 3 ? (233472 / 3 - 1024) : (ctas == 2 ? (233472 / 2 - 1024) : kMaxSmemWs);
  static const int ob_env = getenv("FDL_WS_OB") ? atoi(getenv("FDL_WS_OB")) : 0;
  const int OB0 = ob_env == 1 || ob_env == 2 ? ob_env : (ctas == 3 ? 1 : 2);
  const int G0 = G;
  // Preferred: the padded tile and a refill that trails its tile by one epilogue (NS >= G + 2).  With two epilogue teams a block that
  // does not fit that way (64 -> 64 at 24 x 24: the landmark net) may still run pipelined with the plain pixel stride and two stages.
  static const int tight_env = getenv("FDL_WS_TIGHT") ? atoi(getenv("FDL_WS_TIGHT")) : 1;
  // Last resort for the widest outputs (64 -> 128, 96 -> 96): no output staging at all, the epilogue stores to global memory (OB = 0).
  for (int attempt = 0; attempt < (teams == 2 && tight_env ? 4 : 1); ++attempt) {
    const int pad = (attempt == 0 || attempt == 2) ? in_pad : 0;
    const int OB = attempt >= 2 ? 0 : OB0;
    for (G = G0; G >= 1; --G) {
      const int ndwg = 32 * Q / ipt;
      if (epi + G * ndwg > (ctas == 3 ? 224 : (ctas == 2 ? 320 : kMaxThreads))) continue;
      const int ns_min = attempt >= 1 ? 2 : (ctas == 3 ? 2 : G + 2);
      for (int NS = (G + 3 < kMaxStages ? G + 3 : kMaxStages); NS >= ns_min; --NS) {
        if (ns_cap && NS > ns_cap && NS > G + 2) continue;
        WsLayout L = ws_layout(C, N, Np, wsplit, NS, G, OB, pad, f16);
        if (L.total <= budget) {
          cfg->G = G; cfg->ipt = ipt; cfg->ndwg = ndwg; cfg->NS = NS; cfg->OB = OB; cfg->threads = epi + G * ndwg; cfg->total = L.total; cfg->teams = teams;
          cfg->ctas = ctas; cfg->in_pad = pad; cfg->f16 = f16;
          return true;
        }
      }
    }
  }
  return false;
}

}  // namespace

cudaError_t block_ws_init() {
  cudaError_t e = cudaFuncSetAttribute(block_ws_kernel<512, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<512, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<320, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<512, 1, true, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<512, 1, false, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<224, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<320, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<288, 2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(block_ws_kernel<320, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(block_ws_kernel<320, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemWs);
}

bool block_ws_supported(const Step& s) {
  if (s.kind != STEP_BLOCK || s.w_umma < 0 || s.stride != 1) return false;
  const int C = s.in.C, N = s.out.C;
  if (C % 8 != 0 || N % 4 != 0 || C < 16 || C > 96) return false;
  if (s.Np > 128 || s.out.H < TH || s.out.W < TW) return false;
  if (s.pad_t != 1 || s.pad_l != 1 || s.in.H != s.out.H || s.in.W != s.out.W) return false;
  if (s.in.offset != 0 || s.out.offset != 0 || s.in.batch_stride != (int64_t)s.in.H * s.in.W * C ||
      s.out.batch_stride != (int64_t)s.out.H * s.out.W * N)
    return false;
  if (s.skip.tensor >= 0 && (s.skip_c % 4 != 0 || s.skip.offset != 0)) return false;
  WsCfg cfg;
  const int f16 = (ws_f16_enabled() && s.w_f16 >= 0) ? 1 : 0;
  if (!pick_cfg(C, N, s.Np, f16 ? s.wsplit16 : s.wsplit, f16, &cfg)) return false;
  // a MAX_POOL 2x2 residual from another (twice as large) map: the two-team epilogue reads it with per-thread loads; the others do not
  if (s.skip.tensor >= 0 && s.skip_pool &&
      !(cfg.teams == 2 && s.skip.tensor != s.in.tensor && s.skip.H == 2 * s.out.H && s.skip.W == 2 * s.out.W && s.skip.batch_stride == (int64_t)s.skip.H * s.skip.W * s.skip_c))
    return false;
  return true;
}

cudaError_t launch_block_ws(const BlockTcLaunch& l, cudaStream_t stream) {
  BlockTcArgs a = l.args;
  WsCfg cfg;
  a.f16 = (ws_f16_enabled() && a.w_f16 != nullptr && l.bias_host != nullptr) ? 1 : 0;
  if (!pick_cfg(a.C, a.N, a.Np, a.f16 ? a.wsplit16 : a.wsplit, a.f16, &cfg)) return cudaErrorInvalidConfiguration;
  a.stages = cfg.NS; a.groups = cfg.G; a.dw_threads = cfg.ndwg; a.out_bufs = cfg.OB; a.in_pad = cfg.in_pad;
  a.out_direct = l.out;
  a.pad = 1;
  for (int i = 0; i < 128; ++i) a.alpha_c[i] = (l.alpha_host && i < a.N) ? l.alpha_host[i] : 0.f;
  for (int i = 0; i < 128; ++i) a.bias_c[i] = (a.f16 && i < a.N) ? l.bias_host[i] : 0.f;
  CUtensorMap tm_in, tm_out;
  if (!encode_nhwc(&tm_in, l.in, a.B, a.H, a.W, a.C, (long long)a.H * a.W * a.C, ITH, ITW, cfg.in_pad ? ((a.C / 4) | 1) * 4 : a.C)) return cudaErrorInvalidValue;
  if (!encode_nhwc(&tm_out, l.out, a.B, a.H, a.W, a.N, (long long)a.H * a.W * a.N, TH, TW, ((a.N / 4) | 1) * 4)) return cudaErrorInvalidValue;
  CUtensorMap tm_skip = tm_in;                  // (a valid map when the residual does not travel by TMA)
  a.skip_tma = 0;
  if (cfg.teams == 2 && cfg.OB == 2 && a.skip_mode == 2 && a.skip != nullptr && a.skip_bstride == (long long)a.H * a.W * a.skip_c && a.skip_c % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a.skip) & 15) == 0 && encode_nhwc(&tm_skip, a.skip, a.B, a.H, a.W, a.skip_c, a.skip_bstride, TH, TW, ((a.N / 4) | 1) * 4))
    a.skip_tma = 1;
  a.tiles_x = (a.W + TW - 1) / TW;
  a.tiles_y = (a.H + TH - 1) / TH;
  a.acc_cols = a.Np <= 32 ? 32 : (a.Np <= 64 ? 64 : 128);
  {
    const int need = (cfg.teams == 2 ? kMaxAcc : (cfg.G > 2 ? cfg.G : 2)) * a.acc_cols;
    a.tmem_cols = 32;
    while (a.tmem_cols < need) a.tmem_cols *= 2;
  }
  const int ntiles = a.B * a.tiles_x * a.tiles_y;
  int grid = persist_sms() * cfg.ctas;
  if (grid > ntiles) grid = ntiles;
  cudaError_t e;
  static const int narrow_env = getenv("FDL_WS_NARROW") ? atoi(getenv("FDL_WS_NARROW")) : 1;
  const bool narrow = narrow_env && cfg.ctas == 2 && a.Np <= 32 && a.skip_mode == 1 && cfg.OB == 2;
  static const int mmaw_env = getenv("FDL_WS_MMA_WARP") ? atoi(getenv("FDL_WS_MMA_WARP")) : 1;
  if (narrow && a.f16 && mmaw_env && cfg.threads + kMmaThreads <= 288)
    e = launch_pdl(block_ws_kernel<288, 2, true, true, true>, dim3(grid), dim3(cfg.threads + kMmaThreads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (narrow && a.f16) e = launch_pdl(block_ws_kernel<320, 2, true, true>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (narrow) e = launch_pdl(block_ws_kernel<320, 2, false, true>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (cfg.ctas == 3) e = launch_pdl(block_ws_kernel<224, 3, true>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (cfg.ctas == 2 && a.f16) e = launch_pdl(block_ws_kernel<320, 2, true>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (cfg.ctas == 2) e = launch_pdl(block_ws_kernel<320, 2, false>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (cfg.teams == 2 && a.f16) e = launch_pdl(block_ws_kernel<512, 1, true, false, false, 2>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (cfg.teams == 2) e = launch_pdl(block_ws_kernel<512, 1, false, false, false, 2>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else if (a.f16) e = launch_pdl(block_ws_kernel<512, 1, true>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  else e = launch_pdl(block_ws_kernel<512, 1, false>, dim3(grid), dim3(cfg.threads), (size_t)cfg.total, stream, tm_in, tm_out, tm_skip, a);
  count_launch();
  static const bool verbose = getenv("FDL_WS_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "[ws C=%d N=%d %dx%d G=%d NS=%d ctas=%d thr=%d smem=%d in_pad=%d f16=%d teams=%d]\n", a.C, a.N, a.H, a.W, cfg.G, cfg.NS, cfg.ctas,
            cfg.threads, cfg.total, cfg.in_pad, cfg.f16, cfg.teams);
  return e;
}

}  // namespace fdl
