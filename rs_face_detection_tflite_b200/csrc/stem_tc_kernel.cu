// stem_tc_kernel.cu -- the RGB stem convolution (CONV_2D 5x5 or 3x3, stride 2, Cin = 3, Cout <= 32, + RELU / PRELU) on the tensor
// cores (sm_100a: TMA + tcgen05 + TMEM), warp-specialised.
//
// The FFMA stem (stem_kernel.cu) is bound by the CUDA cores: 75 x 24 multiply-adds per output pixel, 506 us for 256 frames of the
// 256 x 256 detector against ~100 us of HBM traffic.  Here the contraction runs as a GEMM [128 pixels x K'] x [K' x Np]:
//   * one kernel ROW of the window is 3 * KW = 15 (9) CONTIGUOUS floats of the NHWC input row, so the im2col row of an output
//     pixel is KH contiguous segments; each segment is padded to 16 and becomes one K = 16 step of tcgen05.mma kind::f16:
//     K' = 16 KH, two 8-value planes per step,
//   * fp32 fidelity by operand splitting: A = f16 hi + f16 lo planes, W = f16 hi (+ f16 lo when the weights are not f16-exact):
//     2 or 3 MMAs per step, accumulator in TMEM,
//   * the CUDA cores only split and lay out the operand: ~35 instructions per 8 values instead of 24 x 8 multiply-adds.
// Roles of one CTA (two CTAs per SM):
//   warps 0-3, 4-7  two epilogue teams, alternating tiles: TMEM -> + bias, activation -> staging tile -> ONE TMA store per tile
//   warps 8-15      builders: TMA'd input patch ((TH-1)*2+KH rows x ((TW-1)*2+KW)*3 floats) -> A planes, one kernel row (= one K step)
//                   at a time, each with its own "full" barrier: the MMAs of row ky run while row ky + 1 is being laid out
//   warp 16         refills the input ring (TMA loads) and issues the MMAs (one elected lane, operands on the uniform datapath)
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

namespace {

constexpr int TH = 8, TW = 16;                 // output tile: 128 pixels == UMMA M
constexpr int kEpiThreads = 256, kBuildThreads = 256, kThreads = kEpiThreads + kBuildThreads + 32;
constexpr int NS = 3;                          // input patches in flight
constexpr int kPlane = TH * TW * 16 + 16;      // one 8-value plane of A: 128 rows x 16 B (+ 16 B of bank skew)

struct StemTcArgs {
  const float* w = nullptr;      // [K4][Npad] fp32, k = (ky * KW + kx) * 3 + c
  const float* bias = nullptr;
  const float* alpha = nullptr;
  int N = 0, Npad = 0, Np = 32, act = 0, wsplit = 1;   // Np: accumulator columns (32 or 64)
  int KH = 5, KW = 5, pad_t = 0, pad_l = 0;
  int B = 0, tiles_x = 0, tiles_y = 0;
  int patch_rows = 0, patch_floats = 0;    // shared-memory patch: rows x floats per row (a multiple of 4)
  int shift = 0;                           // floats the patch starts before the window: the TMA box must start on a 16-byte boundary of the row
  const int* n_active = nullptr;
};

struct Layout { int bias, w, in0, in_stage, a0, out0, out_stage, total; };
__host__ __device__ inline int align_up_s(int v, int a) { return (v + a - 1) / a * a; }
__host__ __device__ inline Layout layout(int KH, int patch_rows, int patch_floats, int N, int Np) {
  Layout L;
  int off = 128;                               // barriers + tmem slot
  L.bias = off; off += Np * 4;
  L.w = off; off += 2 * (2 * KH) * Np * 16;    // hi planes then lo planes: [2 KH][Np][8 halves] each
  off = align_up_s(off, 128);
  L.in_stage = align_up_s(patch_rows * patch_floats * 4, 128);
  L.in0 = off; off += NS * L.in_stage;
  L.a0 = off; off += align_up_s(2 * (2 * KH) * kPlane, 128);
  L.out_stage = align_up_s(TH * TW * (((N >> 2) | 1) << 2) * 4, 128);
  L.out0 = off; off += 2 * L.out_stage;
  L.total = align_up_s(off, 128);
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

template <int KH>
__global__ void __launch_bounds__(kThreads, 2) stem_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                               const StemTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int Np = a.Np;
  const Layout L = layout(KH, a.patch_rows, a.patch_floats, a.N, Np);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem);       // [NS]  patch landed
  uint64_t* a_full = in_full + NS;                             // [KH]  the two A planes of kernel row ky written by every builder
  uint64_t* a_empty = a_full + KH;                             //       the MMAs have read the A planes
  uint64_t* acc_full = a_empty + 1;                            // [2]   accumulator complete
  uint64_t* acc_empty = acc_full + 2;                          // [2]   accumulator drained by its epilogue team
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint32_t* w_inexact = tmem_slot + 1;                         // some weight is not an f16: the W_lo pass is needed
  constexpr int KQ = 2 * KH;                                   // planes per operand half (hi / lo)
  const int KW3 = a.KW * 3;

  // ---- prologue: nothing here depends on the previous launch (PDL) ----
  if (tid == 0) {
    ptx::prefetch_tmap(&tm_in);
    ptx::prefetch_tmap(&tm_out);
    for (int s = 0; s < NS; ++s) ptx::mbar_init(&in_full[s], 1);
    for (int k = 0; k < KH; ++k) ptx::mbar_init(&a_full[k], kBuildThreads);
    *w_inexact = 0;
    ptx::mbar_init(a_empty, 1);
    for (int t = 0; t < 2; ++t) { ptx::mbar_init(&acc_full[t], 1); ptx::mbar_init(&acc_empty[t], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, (uint32_t)(2 * Np));
  __syncthreads();
  {
    // weights: fp32 [k][n] in global memory -> f16 hi / lo planes in the K' order of the A operand (k' = 16 ky + j, j = 3 kx + c < 3 KW)
    uint16_t* w_hi = reinterpret_cast<uint16_t*>(smem + L.w);
    uint16_t* w_lo = w_hi + KQ * Np * 8;
    for (int i = tid; i < KQ * Np * 8; i += kThreads) {
      const int kq = i / (Np * 8), r = i - kq * Np * 8, n = r >> 3, e = r & 7;
      const int ky = kq >> 1, j = (kq & 1) * 8 + e;
      float v = 0.f;
      if (j < KW3 && n < a.N) v = __ldg(a.w + (long long)(ky * KW3 + j) * a.Npad + n);
      const uint16_t h = f2h(v);
      w_hi[i] = h;
      const uint16_t l = f2h(v - h2f(h));
      w_lo[i] = l;
      if (l & 0x7fff) atomicOr(w_inexact, 1u);
    }
  }
  if (tid < Np) reinterpret_cast<float*>(smem + L.bias)[tid] = tid < a.N ? __ldg(a.bias + tid) : 0.f;
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const bool wsplit2 = __shfl_sync(0xffffffffu, *w_inexact, 0) != 0;
  pdl_launch_dependents();
  pdl_wait();
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int ntiles = nb * tiles_per_img;
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t in_bytes = (uint32_t)(a.patch_rows * a.patch_floats * 4);

  if (my_tiles == 0) {
    // nothing to do
  } else if (warp_u < 8) {
    // ================= epilogue teams: team e takes the CTA's tiles e, e + 2, ... and owns accumulator e / staging buffer e =================
    const int e = warp_u >> 2, p = tid & 127;                  // TMEM lane == pixel of the tile
    const int NPf = ((a.N >> 2) | 1) << 2;                     // staging pixel stride (floats): an odd number of quads
    float* s_o = reinterpret_cast<float*>(smem + L.out0 + e * L.out_stage) + p * NPf;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp_u & 3) * 32) << 16) + (uint32_t)(e * Np);
    const bool leader = p == 0;
    const int bar_id = 1 + e;
    const float* s_bias = reinterpret_cast<const float*>(smem + L.bias);
    for (int it = e, k = 0; it < my_tiles; it += 2, ++k) {
      ptx::mbar_wait(&acc_full[e], (uint32_t)(k & 1));
      ptx::tc_fence_after_sync();
      uint32_t r0[16], r1[16];
      ptx::tmem_ld16_issue(taddr, r0);                         // (in flight across the hand-over of the staging buffer)
      ptx::tmem_ld16_issue(taddr + 16u, r1);
      if (leader) ptx::tma_store_wait_read0();                 // this team's previous store has read the staging buffer
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      for (int c0 = 0; c0 < Np; c0 += 32) {
        if (c0 > 0) {
          ptx::tmem_ld16_issue(taddr + (uint32_t)c0, r0);
          ptx::tmem_ld16_issue(taddr + (uint32_t)c0 + 16u, r1);
        }
        ptx::tmem_ld_wait16(r0);
        ptx::tmem_ld_wait16(r1);
        if (c0 + 32 >= Np) {                                   // accumulator drained
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[e]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = c0 + 16 * h + 4 * j;
            if (n < a.N) {
              const uint32_t* r = h ? r1 : r0;
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n);
              float4 o = make_float4(__uint_as_float(r[4 * j]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y, __uint_as_float(r[4 * j + 2]) + b4.z,
                                     __uint_as_float(r[4 * j + 3]) + b4.w);
              if (a.act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              else if (a.act == ACT_PRELU) {
                const float4 al = __ldg(reinterpret_cast<const float4*>(a.alpha + n));
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
        const int b = tile / tiles_per_img, rr = tile - b * tiles_per_img, ty = rr / a.tiles_x, tx = rr - ty * a.tiles_x;
        ptx::tma_store_4d(&tm_out, smem + L.out0 + e * L.out_stage, 0, tx * TW, ty * TH, b);
        ptx::tma_store_commit();
      }
    }
    if (leader) ptx::tma_store_wait_all0();
  } else if (warp_u < (kEpiThreads + kBuildThreads) / 32) {
    // ================= builders: patch -> A planes (hi, lo) =================
    const int bt = tid - kEpiThreads, p = bt & 127, half = bt >> 7, py = p >> 4, px = p & 15;   // one pixel, the first or the second 8 values of a kernel row
    uint8_t* s_a = smem + L.a0;
    const int pf = a.patch_floats, shift = a.shift, j0 = half * 8;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NS;
      ptx::mbar_wait(&in_full[s], (uint32_t)((it / NS) & 1));
      if (it > 0) ptx::mbar_wait(a_empty, (uint32_t)((it - 1) & 1));
      const float* s_in = reinterpret_cast<const float*>(smem + L.in0 + s * L.in_stage);
      // all the loads of the tile first (the fence / arrive pairs below are compiler barriers: nothing would be hoisted over them)
      float vv[KH][8];
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const float* src = s_in + (2 * py + ky) * pf + 6 * px + j0 + shift;
        if (shift & 1) {                                                 // odd start: five aligned 8-byte loads around the eight values
          float2 t[5];
#pragma unroll
          for (int u = 0; u < 5; ++u) t[u] = *reinterpret_cast<const float2*>(src - 1 + 2 * u);
#pragma unroll
          for (int u = 0; u < 4; ++u) { vv[ky][2 * u] = t[u].y; vv[ky][2 * u + 1] = t[u + 1].x; }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 t = *reinterpret_cast<const float2*>(src + 2 * u);
            vv[ky][2 * u] = t.x; vv[ky][2 * u + 1] = t.y;
          }
        }
      }
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int kq = 2 * ky + half;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (j0 + u >= KW3) ? 0.f : vv[ky][u];   // the padding of the segment to 16 (the patch holds the next pixels there)
        // hi = the value cut to 11 significant bits (a mask: exactly an f16 in the normal range, so its conversion does not round and
        // nothing has to be converted back), lo = the rest, rounded to f16: hi + lo carries >= 21 bits of the value.  (Below the f16
        // normal range, |v| < 6.1e-5, the conversion of hi rounds by at most 3e-8 absolute: the pixels are O(1).)
        float hf[8], lf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { hf[u] = __uint_as_float(__float_as_uint(v[u]) & 0xffffe000u); lf[u] = v[u] - hf[u]; }
        uint4 hi, lo;
        hi.x = pack_f16x2(hf[0], hf[1]); hi.y = pack_f16x2(hf[2], hf[3]); hi.z = pack_f16x2(hf[4], hf[5]); hi.w = pack_f16x2(hf[6], hf[7]);
        lo.x = pack_f16x2(lf[0], lf[1]); lo.y = pack_f16x2(lf[2], lf[3]); lo.z = pack_f16x2(lf[4], lf[5]); lo.w = pack_f16x2(lf[6], lf[7]);
        *reinterpret_cast<uint4*>(s_a + kq * kPlane + p * 16) = hi;
        *reinterpret_cast<uint4*>(s_a + (KQ + kq) * kPlane + p * 16) = lo;
        ptx::fence_proxy_async_smem();
        mbar_arrive(&a_full[ky]);
      }
    }
  } else {
    // ================= warp 16: input ring + MMA issue (the whole warp, converged; one elected lane issues) =================
    auto issue_load = [&](int it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x, s = it % NS;
      const int b = tile / tiles_per_img, rr = tile - b * tiles_per_img, ty = rr / a.tiles_x, tx = rr - ty * a.tiles_x;
      ptx::mbar_arrive_expect_tx(&in_full[s], in_bytes);
      ptx::tma_load_4d(smem + L.in0 + s * L.in_stage, &tm_in, &in_full[s], (tx * TW * 2 - a.pad_l) * 3 - a.shift, 0, ty * TH * 2 - a.pad_t, b);
    };
    if (lane == 0)
      for (int it = 0; it < NS && it < my_tiles; ++it) issue_load(it);
    __syncwarp();
    const uint32_t idesc = ptx::umma_idesc_f16(128, Np);
    const uint32_t a_hi = ptx::smem_u32(smem + L.a0), a_lo = a_hi + (uint32_t)(KQ * kPlane);
    const uint32_t w_hi = ptx::smem_u32(smem + L.w), w_lo = w_hi + (uint32_t)(KQ * Np * 16);
    const uint32_t lbo_w = (uint32_t)Np * 16u;
    for (int it = 0; it < my_tiles; ++it) {
      const int t = it & 1;
      const uint32_t d_tmem = tmem_base + (uint32_t)(t * Np);
#pragma unroll
      for (int ks = 0; ks < KH; ++ks) {
        ptx::mbar_wait(&a_full[ks], (uint32_t)(it & 1));
        if (ks == 0 && it >= 2) ptx::mbar_wait(&acc_empty[t], (uint32_t)(((it >> 1) - 1) & 1));
        if (ks == KH - 1) {
          // every builder is done with the patch of this tile: its stage takes the tile NS places ahead
          if (lane == 0 && it + NS < my_tiles) issue_load(it + NS);
          __syncwarp();
        }
        ptx::tc_fence_after_sync();
        const uint64_t dah = ptx::umma_desc_kmajor(a_hi + (uint32_t)(2 * ks * kPlane), kPlane, 128);
        const uint64_t dal = ptx::umma_desc_kmajor(a_lo + (uint32_t)(2 * ks * kPlane), kPlane, 128);
        const uint64_t dbh = ptx::umma_desc_kmajor(w_hi + (uint32_t)(2 * ks) * lbo_w, lbo_w, 128);
        ptx::mma_f16_elect(d_tmem, dah, dbh, idesc, ks ? 1u : 0u);
        ptx::mma_f16_elect(d_tmem, dal, dbh, idesc, 1u);
        if (wsplit2) {
          const uint64_t dbl = ptx::umma_desc_kmajor(w_lo + (uint32_t)(2 * ks) * lbo_w, lbo_w, 128);
          ptx::mma_f16_elect(d_tmem, dah, dbl, idesc, 1u);
        }
      }
      ptx::mma_commit_elect(a_empty);
      ptx::mma_commit_elect(&acc_full[t]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)(2 * Np));
}

struct Cfg { int patch_rows, patch_floats, tiles_x, tiles_y, total; };
Cfg cfg_of(const ConvArgs& a) {
  Cfg c;
  c.patch_rows = (TH - 1) * 2 + a.kh;
  // the builders read 16 floats per kernel row from float 6 px + shift on (shift <= 2): the last pixel's segment ends within 108 floats
  c.patch_floats = 108;
  c.tiles_x = a.out.W / TW;
  c.tiles_y = a.out.H / TH;
  c.total = layout(a.kh, c.patch_rows, c.patch_floats, a.N, a.N <= 32 ? 32 : 64).total;
  return c;
}

}  // namespace

cudaError_t stem_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e;
}

bool stem_tc_supported(const ConvArgs& a) {
  static const bool on = [] { const char* e = getenv("FDL_STEM_TC"); return e ? atoi(e) != 0 : true; }();
  if (!on || !a.mma) return false;
  if (a.mode != 0 || a.in.C != 3 || a.stride != 2 || a.kh != a.kw || (a.kh != 3 && a.kh != 5) || a.has_skip) return false;
  if (a.N % 4 != 0 || a.N > 64 || a.N < 8) return false;
  if (a.out.H % TH != 0 || a.out.W % TW != 0) return false;
  if (a.out.bstride != (long long)a.out.H * a.out.W * a.N || a.in.bstride != (long long)a.in.H * a.in.W * 3) return false;
  if ((a.in.W * 3) % 4 != 0 || (reinterpret_cast<uintptr_t>(a.in.p) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.out.p) & 15) != 0) return false;
  if (a.pad_t < 0 || a.pad_t > 2 || a.pad_l < 0 || a.pad_l > 2) return false;
  return cfg_of(a).total <= 227 * 1024;
}

cudaError_t launch_stem_tc(const ConvArgs& a, cudaStream_t stream) {
  const Cfg c = cfg_of(a);
  StemTcArgs k;
  k.w = a.w; k.bias = a.bias; k.alpha = a.alpha; k.N = a.N; k.Npad = a.Npad; k.Np = a.N <= 32 ? 32 : 64; k.act = a.act; k.wsplit = 2;
  k.KH = a.kh; k.KW = a.kw; k.pad_t = a.pad_t; k.pad_l = a.pad_l; k.B = a.B; k.tiles_x = c.tiles_x; k.tiles_y = c.tiles_y;
  k.patch_rows = c.patch_rows; k.patch_floats = c.patch_floats; k.n_active = a.n_active;
  k.shift = (4 - (a.pad_l * 3) % 4) % 4;     // (2 tx TW - pad_l) * 3 - shift is a multiple of 4 floats
  CUtensorMap tm_in, tm_out;
  // the input as [B][H][1][W * 3]: the innermost TMA dimension is a whole image row, so a box of 108 floats is 36 pixels of 3 channels
  if (!encode_nhwc(&tm_in, a.in.p, a.B, a.in.H, 1, a.in.W * 3, a.in.bstride, c.patch_rows, 1, c.patch_floats)) return cudaErrorInvalidValue;
  if (!encode_nhwc(&tm_out, a.out.p, a.B, a.out.H, a.out.W, a.N, a.out.bstride, TH, TW, ((a.N / 4) | 1) * 4)) return cudaErrorInvalidValue;
  const int ntiles = a.B * c.tiles_x * c.tiles_y;
  if (ntiles == 0) return cudaSuccess;
  int grid = persist_sms() * (c.total <= 113 * 1024 ? 2 : 1);       // two CTAs per SM when they fit
  if (grid > ntiles) grid = ntiles;
  cudaError_t e;
  if (a.kh == 5) e = launch_pdl(stem_tc_kernel<5>, dim3(grid), dim3(kThreads), (size_t)c.total, stream, tm_in, tm_out, k);
  else e = launch_pdl(stem_tc_kernel<3>, dim3(grid), dim3(kThreads), (size_t)c.total, stream, tm_in, tm_out, k);
  count_launch();
  return e;
}

}  // namespace fdl
