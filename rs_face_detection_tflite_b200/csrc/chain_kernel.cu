// chain_kernel.cu -- the tail-chain kernel (see chain.h): one persistent CTA per SM walks the chain's program for a group of
// items with every activation in shared memory; the pointwise / patch convolutions run on the tensor cores (tcgen05.mma
// kind::f16, operands split as f16 hi + f16 lo, accumulator in TMEM), weights arrive through a cp.async.bulk ring fed by a
// producer warp one layer ahead.  Only the chain's inputs and outputs touch global memory.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "chain.h"
#include "chain_kernel.cuh"
#include "pdl.h"
#include "plan.h"
#include "sm100_ptx.cuh"

namespace fdl {

void count_launch();

// FDL_WS_TRACE (variant build only): globaltimer stamps of CTA 0's first two groups -- [group][0] = start, [group][1 + op] = the op's
// closing barrier passed.
#ifdef FDL_WS_TRACE
__device__ unsigned long long g_chain_trace[2][kChainMaxOps + 1];
__device__ unsigned long long g_chain_trace2[kChainMaxOps][4];   // inside a GEMM of group 0: MMAs issued, accumulator seen, epilogue done
#define CH_T2(grp_, oi_, k_)                                                         \
  do {                                                                               \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (grp_) == 0) {                        \
      unsigned long long t_;                                                         \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                          \
      g_chain_trace2[(oi_)][(k_)] = t_;                                              \
    }                                                                                \
  } while (0)
#define CH_T(grp_, slot_)                                                            \
  do {                                                                               \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (grp_) < 2) {                         \
      unsigned long long t_;                                                         \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                          \
      g_chain_trace[(grp_)][(slot_)] = t_;                                           \
    }                                                                                \
  } while (0)
cudaError_t chain_trace_read(unsigned long long* out, int n) {
  const int n1 = 2 * (kChainMaxOps + 1), n2 = kChainMaxOps * 4;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_chain_trace, sizeof(unsigned long long) * (size_t)(n < n1 ? n : n1));
  if (e == cudaSuccess && n >= n1 + n2) e = cudaMemcpyFromSymbol(out + n1, g_chain_trace2, sizeof(unsigned long long) * (size_t)n2);
  return e;
}
#else
#define CH_T(grp_, slot_) do { } while (0)
#define CH_T2(grp_, oi_, k_) do { } while (0)
cudaError_t chain_trace_read(unsigned long long*, int) { return cudaErrorNotSupported; }
#endif

namespace {

constexpr int kOffOps = 256;                // the program, copied from the kernel parameters: a phase reads its op with 8 LDS.128
constexpr int kOffArena = kOffOps + kChainMaxOps * (int)sizeof(ChainOp);
constexpr int kOffW = kOffArena + kChainArena;
constexpr int kOffPar = kOffW + 2 * kChainWSlot;
constexpr int kChainSmem = kOffPar + 2 * kChainParSlot;
static_assert(kOffArena % 128 == 0 && kChainSmem <= 227 * 1024, "chain kernel shared-memory map");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kChainThreads) : "memory"); }

__device__ __forceinline__ uint32_t pack_f16x2(float c0, float c1) {   // c0 in the low half; round to nearest even, saturating
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(c1), "f"(c0));
  return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t d) {
  float2 r;
  asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}\n" : "=f"(r.x), "=f"(r.y) : "r"(d));
  return r;
}
// four floats -> (hi halves, lo halves)
__device__ __forceinline__ void split4(const float4& v, uint2* hi, uint2* lo) {
  hi->x = pack_f16x2(v.x, v.y); hi->y = pack_f16x2(v.z, v.w);
  const float2 h01 = unpack_f16x2(hi->x), h23 = unpack_f16x2(hi->y);
  lo->x = pack_f16x2(v.x - h01.x, v.y - h01.y); lo->y = pack_f16x2(v.z - h23.x, v.w - h23.y);
}
__device__ __forceinline__ float4 join4(const uint2& hi, const uint2& lo) {
  const float2 h01 = unpack_f16x2(hi.x), h23 = unpack_f16x2(hi.y), l01 = unpack_f16x2(lo.x), l23 = unpack_f16x2(lo.y);
  return make_float4(h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y);
}
__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fma4c(float4& a, const float4& x, const float4& w) {
  a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// channels [4q, 4q + 4) of row `row` of a tensor in either format
__device__ __forceinline__ float4 read_quad(const uint8_t* arena, const ChainTensor& t, int row, int q) {
  if (t.fmt == CH_F32) return *reinterpret_cast<const float4*>(arena + t.off + ((size_t)row * t.pl + 4 * q) * 4);
  const uint8_t* ph = arena + t.off + (q >> 1) * t.pl + row * 16 + (q & 1) * 8;
  const uint2 hi = *reinterpret_cast<const uint2*>(ph), lo = *reinterpret_cast<const uint2*>(ph + (t.C >> 3) * t.pl);
  return join4(hi, lo);
}
__device__ __forceinline__ void write_quad(uint8_t* arena, const ChainTensor& t, int row, int q, const float4& v) {
  if (t.fmt == CH_F32) { *reinterpret_cast<float4*>(arena + t.off + ((size_t)row * t.pl + 4 * q) * 4) = v; return; }
  uint2 hi, lo;
  split4(v, &hi, &lo);
  uint8_t* ph = arena + t.off + (q >> 1) * t.pl + row * 16 + (q & 1) * 8;
  *reinterpret_cast<uint2*>(ph) = hi;
  *reinterpret_cast<uint2*>(ph + (t.C >> 3) * t.pl) = lo;
}

// 16 accumulator columns [c0, c0 + 16) of one row: + bias, + residual, activation, written in the output's format.  The formats are
// compile-time (kOutP16: the output is P16, else F32; kSkip: 0 none, 1 F32, 2 P16) so that plane rows move as whole 16-byte chunks.
template <bool kOutP16, int kSkip>
__device__ __forceinline__ void epilogue16(uint8_t* arena, const float (&v)[16], int row, int c0, const float* bias, const float* alpha, int act,
                                           const ChainTensor& out, const ChainTensor& skip, int skip_c) {
  float r[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + 4 * j);
    r[4 * j] = v[4 * j] + b4.x; r[4 * j + 1] = v[4 * j + 1] + b4.y; r[4 * j + 2] = v[4 * j + 2] + b4.z; r[4 * j + 3] = v[4 * j + 3] + b4.w;
  }
  if (kSkip == 1) {
    const float* sp = reinterpret_cast<const float*>(arena + skip.off) + (size_t)row * skip.pl + c0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c0 + 4 * j < skip_c) {
        const float4 s = *reinterpret_cast<const float4*>(sp + 4 * j);
        r[4 * j] += s.x; r[4 * j + 1] += s.y; r[4 * j + 2] += s.z; r[4 * j + 3] += s.w;
      }
  } else if (kSkip == 2) {
    const uint8_t* ph = arena + skip.off + (c0 >> 3) * skip.pl + row * 16;
    const int lo_off = (skip.C >> 3) * skip.pl;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (c0 + 8 * h < skip_c) {
        const uint4 hi = *reinterpret_cast<const uint4*>(ph + h * skip.pl), lo = *reinterpret_cast<const uint4*>(ph + h * skip.pl + lo_off);
        const float4 s0 = join4(make_uint2(hi.x, hi.y), make_uint2(lo.x, lo.y)), s1 = join4(make_uint2(hi.z, hi.w), make_uint2(lo.z, lo.w));
        r[8 * h] += s0.x; r[8 * h + 1] += s0.y; r[8 * h + 2] += s0.z; r[8 * h + 3] += s0.w;
        r[8 * h + 4] += s1.x; r[8 * h + 5] += s1.y; r[8 * h + 6] += s1.z; r[8 * h + 7] += s1.w;
      }
  }
  if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = fmaxf(r[j], 0.f);
  } else if (act == ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 al = *reinterpret_cast<const float4*>(alpha + c0 + 4 * j);
      r[4 * j] = r[4 * j] >= 0.f ? r[4 * j] : r[4 * j] * al.x; r[4 * j + 1] = r[4 * j + 1] >= 0.f ? r[4 * j + 1] : r[4 * j + 1] * al.y;
      r[4 * j + 2] = r[4 * j + 2] >= 0.f ? r[4 * j + 2] : r[4 * j + 2] * al.z; r[4 * j + 3] = r[4 * j + 3] >= 0.f ? r[4 * j + 3] : r[4 * j + 3] * al.w;
    }
  }
  if (kOutP16) {
    uint8_t* ph = arena + out.off + (c0 >> 3) * out.pl + row * 16;
    const int lo_off = (out.C >> 3) * out.pl;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint2 h0, l0, h1, l1;
      split4(make_float4(r[8 * h], r[8 * h + 1], r[8 * h + 2], r[8 * h + 3]), &h0, &l0);
      split4(make_float4(r[8 * h + 4], r[8 * h + 5], r[8 * h + 6], r[8 * h + 7]), &h1, &l1);
      *reinterpret_cast<uint4*>(ph + h * out.pl) = make_uint4(h0.x, h0.y, h1.x, h1.y);
      *reinterpret_cast<uint4*>(ph + h * out.pl + lo_off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
    }
  } else {
    float* op = reinterpret_cast<float*>(arena + out.off) + (size_t)row * out.pl + c0;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(op + 4 * j) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  }
}

__global__ void __launch_bounds__(kChainThreads + 32, 1) chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);      // [2] weight chunk landed
  uint64_t* w_empty = w_full + 2;                            // [2] the MMAs that read the slot are complete
  uint64_t* p_full = w_empty + 2;                            // [2] parameter block landed
  uint64_t* p_empty = p_full + 2;                            // [2] the step that read the block is over
  uint64_t* acc_full = p_empty + 2;                          // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const ChainOp* s_ops = reinterpret_cast<const ChainOp*>(smem + kOffOps);
  uint8_t* arena = smem + kOffArena;
  uint8_t* s_w = smem + kOffW;
  uint8_t* s_par = smem + kOffPar;

  // ---- prologue: nothing here depends on the previous launch (PDL) ----
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); ptx::mbar_init(&p_full[i], 1); ptx::mbar_init(&p_empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 128);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.ops);
    uint4* dst = reinterpret_cast<uint4*>(smem + kOffOps);
    for (int i = tid; i < a.n_ops * (int)(sizeof(ChainOp) / 16); i += kChainThreads + 32) dst[i] = src[i];
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (provably warp-uniform: see mma_f16_elect)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_launch_dependents();
  pdl_wait();
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const int G = a.items;
  const int n_groups = (nb + G - 1) / G;
  const int my_groups = (int)blockIdx.x < n_groups ? (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kChainThreads / 32) {
    // ================= producer: weight chunks and parameter blocks, in program order, as far ahead as the rings allow =================
    if (lane == 0) {
      int wi = 0, pi = 0;
      for (int grp = 0; grp < my_groups; ++grp) {
        for (int l = 0; l < a.n_loads; ++l) {
          const ChainLoad& ld = a.loads[l];
          if (ld.is_par) {
            const int buf = pi & 1;
            ptx::mbar_wait(&p_empty[buf], (uint32_t)(((pi >> 1) & 1) ^ 1));
            ptx::mbar_arrive_expect_tx(&p_full[buf], (uint32_t)ld.bytes);
            ptx::bulk_load_1d(s_par + buf * kChainParSlot, a.weights + ld.w_off, (uint32_t)ld.bytes, &p_full[buf]);
            ++pi;
          } else {
            const int buf = wi & 1;
            ptx::mbar_wait(&w_empty[buf], (uint32_t)(((wi >> 1) & 1) ^ 1));
            ptx::mbar_arrive_expect_tx(&w_full[buf], (uint32_t)ld.bytes);
            ptx::bulk_load_1d(s_w + buf * kChainWSlot, a.weights + ld.w_off, (uint32_t)ld.bytes, &w_full[buf]);
            ++wi;
          }
        }
      }
    }
  } else if (my_groups > 0) {
    // ================= workers =================
    int wi = 0, pi = 0, acc_phase = 0;                         // ring positions: identical in every worker thread
    const uint32_t arena_addr = ptx::smem_u32(arena), w_addr = ptx::smem_u32(s_w);
    float* const g_arena = a.arena;
    for (int grp = 0; grp < my_groups; ++grp) {
      const int item0 = ((int)blockIdx.x + grp * (int)gridDim.x) * G;
      const int nv = min(G, nb - item0);                        // valid items of this group
      CH_T(grp, 0);
      for (int oi = 0; oi < a.n_ops; ++oi) {
        const ChainOp o = s_ops[oi];                           // a register copy (constant-bank reads of the op cost ~0.3 us per phase)
        const int kind = o.kind;
        if (kind == CH_LOAD || kind == CH_STORE) {
          const ChainTensor t = kind == CH_LOAD ? o.out : o.in;
          const int Q = t.C >> 2, hw = t.H * t.W, n = nv * hw * Q;
          const long long bstride = o.g_bstride;
          float* base = g_arena + o.g_buf_offset * (long long)a.B + o.g_offset + (long long)item0 * bstride;
          // the group's items are consecutive in global memory when the tensor is dense: element e of the group is at base + e
          // (bstride == hw * C), so quad `it` is simply the it-th float4
          if (kind == CH_LOAD) {
            for (int it0 = tid; it0 < n; it0 += 8 * kChainThreads) {
              float4 v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * kChainThreads;
                if (it < n) v[u] = *(reinterpret_cast<const float4*>(base) + it);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * kChainThreads;
                if (it < n) write_quad(arena, t, it / Q, it % Q, v[u]);
              }
            }
          } else {
            for (int it = tid; it < n; it += kChainThreads) *(reinterpret_cast<float4*>(base) + it) = read_quad(arena, t, it / Q, it % Q);
          }
        } else if (kind == CH_POOL) {
          const ChainTensor& ti = o.in;
          const ChainTensor& to = o.out;
          const int Q = to.C >> 2, hw = to.H * to.W, n = G * hw * Q;
          for (int it = tid; it < n; it += kChainThreads) {
            const int q = it % Q, row = it / Q, i = row / hw, r = row - i * hw, y = r / to.W, x = r - y * to.W;
            const int s0 = i * ti.H * ti.W + (2 * y) * ti.W + 2 * x;
            const float4 v = max4(max4(read_quad(arena, ti, s0, q), read_quad(arena, ti, s0 + 1, q)),
                                  max4(read_quad(arena, ti, s0 + ti.W, q), read_quad(arena, ti, s0 + ti.W + 1, q)));
            write_quad(arena, to, row, q, v);
          }
        } else if (kind == CH_GATHER) {
          // im2col of a k x k / stride k convolution: whole 16-byte plane rows move, K order = (tap, channel).  A thread keeps one
          // (output row, tap) pair and walks the planes: no divisions in the copy loop.
          const ChainTensor ti = o.in, to = o.out;
          const int k = o.k, hw = to.H * to.W, rows = G * hw, kq_in = ti.C >> 3, kq_out = to.C >> 3, npairs = rows * k * k;
          const int nparts = npairs < kChainThreads ? kChainThreads / npairs : 1;      // threads sharing one pair split the planes
          const int part = tid / npairs;
          if (part < nparts)
            for (int pr = tid - part * npairs; pr < npairs; pr += kChainThreads) {
              const int row = pr % rows, tap = pr / rows;
              const int i = row / hw, r = row - i * hw, y = r / to.W, x = r - y * to.W;
              const int src = i * ti.H * ti.W + (y * k + tap / k) * ti.W + (x * k + tap % k);
              const uint8_t* sp = arena + ti.off + src * 16;
              uint8_t* dp = arena + to.off + tap * kq_in * to.pl + row * 16;
              for (int j = part; j < kq_in; j += nparts) {
                *reinterpret_cast<uint4*>(dp + j * to.pl) = *reinterpret_cast<const uint4*>(sp + j * ti.pl);
                *reinterpret_cast<uint4*>(dp + (kq_out + j) * to.pl) = *reinterpret_cast<const uint4*>(sp + (kq_in + j) * ti.pl);
              }
            }
        } else if (kind == CH_DW) {
          // depthwise 3x3 + bias on the CUDA cores: F32 in, the GEMM's A operand (P16) out
          CH_T2(grp, oi, 3);
          ptx::mbar_wait(&p_full[pi & 1], (uint32_t)((pi >> 1) & 1));
          CH_T2(grp, oi, 0);
          const float* par = reinterpret_cast<const float*>(s_par + (pi & 1) * kChainParSlot);
          const ChainTensor ti = o.in, to = o.out;
          const int C = ti.C, Q = C >> 2, hw = to.H * to.W, n = G * hw * Q, S = o.stride, pad_t = o.pad_t, pad_l = o.pad_l;
          const int Hi = ti.H, Wi = ti.W, Wo = to.W, ipl = ti.pl;
          // rows < 128 and divisors <= 64: floor(n / d) == (n * ceil(2^16 / d)) >> 16
          const unsigned m_hw = (65536u + hw - 1) / hw, m_w = (65536u + Wo - 1) / Wo;
          const int q = tid % Q, row0 = tid / Q, rstep = kChainThreads / Q;   // kChainThreads % Q == 0: the channel quad is fixed per thread
          const float4 bd = *reinterpret_cast<const float4*>(par + 9 * C + 4 * q);
          const float* in_q = reinterpret_cast<const float*>(arena + ti.off) + 4 * q;
          const int nrows = n / Q;
          CH_T2(grp, oi, 1);
          if (S == 1 && pad_t == 1 && pad_l == 1 && Hi == to.H && Wi == Wo) {
            // stride 1: a thread walks one COLUMN of one item down the map with the three rows of its 3 x 3 window in registers --
            // three loads per output instead of nine, no border tests inside the window (the missing neighbours are zeros)
            const int ncols = G * Wo;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* wq = reinterpret_cast<const float4*>(par) + q;     // (weights re-read from shared memory: the register budget is 96)
#define W(k_) wq[(k_) * Q]
            for (int cx = row0; cx < ncols; cx += rstep) {
              const int i = (int)(((unsigned)cx * m_w) >> 16), x = cx - i * Wo;
              const bool vl = x > 0, vr = x + 1 < Wi;
              const float* col = in_q + (size_t)(i * Hi * Wi + x) * ipl;
              // row r is the bottom row of output r - 1's window, the middle row of output r's and the top row of output r + 1's:
              // three running sums, one finished per row
              float4 o_m1 = bd, o_0 = bd, o_p1 = bd;
              for (int r = 0; r < Hi; ++r) {
                const float* rp = col + (size_t)r * Wi * ipl;
                const float4 v0 = vl ? ld4f(rp - ipl) : z4, v1 = ld4f(rp), v2 = vr ? ld4f(rp + ipl) : z4;
                if (r >= 1) {
                  fma4c(o_m1, v0, W(6)); fma4c(o_m1, v1, W(7)); fma4c(o_m1, v2, W(8));
                  write_quad(arena, to, i * hw + (r - 1) * Wo + x, q, o_m1);
                }
                fma4c(o_0, v0, W(3)); fma4c(o_0, v1, W(4)); fma4c(o_0, v2, W(5));
                fma4c(o_p1, v0, W(0)); fma4c(o_p1, v1, W(1)); fma4c(o_p1, v2, W(2));
                o_m1 = o_0; o_0 = o_p1; o_p1 = bd;
              }
              write_quad(arena, to, i * hw + (Hi - 1) * Wo + x, q, o_m1);     // (its bottom row is the zero padding)
            }
#undef W
          } else {
          float4 wd[9];
#pragma unroll
          for (int kk = 0; kk < 9; ++kk) wd[kk] = *reinterpret_cast<const float4*>(par + kk * C + 4 * q);
          for (int row = row0; row < nrows; row += rstep) {
            const int i = (int)(((unsigned)row * m_hw) >> 16), r = row - i * hw, y = (int)(((unsigned)r * m_w) >> 16), x = r - y * Wo;
            const int iy0 = y * S - pad_t, ix0 = x * S - pad_l;
            const float* ip = in_q + (size_t)(i * Hi * Wi + iy0 * Wi + ix0) * ipl;
            float4 acc = bd;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const bool vy = (unsigned)(iy0 + ky) < (unsigned)Hi;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                // (a branch-free form -- clamped addresses, zeroed weights -- was measured slower: 6.1 vs 3.4 us on the 8 x 8 maps)
                if (vy && (unsigned)(ix0 + kx) < (unsigned)Wi) {
                  const float4 v = *reinterpret_cast<const float4*>(ip + (ky * Wi + kx) * ipl);
                  const float4 w = wd[ky * 3 + kx];
                  acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
                }
              }
            }
            write_quad(arena, to, row, q, acc);
          }
          }
        } else {
          // ---- GEMM: [128 rows x K] x [K x Np] -> TMEM, then the epilogue ----
          const int Np = o.Np, K = o.K;
          if (warp_u == 0) {                                     // the whole warp, converged: one elected lane issues
            const uint32_t idesc = ptx::umma_idesc_f16(128, Np);
            const int kc_max = chain_kc_max(Np);
            const uint32_t a_hi = arena_addr + (uint32_t)o.in.off, a_lo = a_hi + (uint32_t)((K >> 3) * o.in.pl);
            const uint32_t lbo_w = (uint32_t)Np * 16u;
            uint32_t accf = 0;
            const uint64_t a_step = (uint64_t)((2u * (uint32_t)o.in.pl) >> 4), b_step = (uint64_t)((2u * lbo_w) >> 4);
            uint64_t dah = ptx::umma_desc_kmajor(a_hi, (uint32_t)o.in.pl, 128), dal = ptx::umma_desc_kmajor(a_lo, (uint32_t)o.in.pl, 128);
            for (int j = 0; j < o.nchunks; ++j) {
              const int k0 = j * kc_max, kc = min(kc_max, K - k0), buf = (wi + j) & 1;
              ptx::mbar_wait(&w_full[buf], (uint32_t)(((wi + j) >> 1) & 1));
              ptx::tc_fence_after_sync();
              const uint32_t b_hi = w_addr + (uint32_t)(buf * kChainWSlot), b_lo = b_hi + (uint32_t)((kc >> 3) * Np * 16);
              uint64_t dbh = ptx::umma_desc_kmajor(b_hi, lbo_w, 128), dbl = ptx::umma_desc_kmajor(b_lo, lbo_w, 128);
              for (int ks = 0; ks < kc; ks += 16, dah += a_step, dal += a_step, dbh += b_step, dbl += b_step) {
                ptx::mma_f16_elect(tmem_base, dah, dbh, idesc, accf);
                accf = 1;
                ptx::mma_f16_elect(tmem_base, dal, dbh, idesc, 1u);
                ptx::mma_f16_elect(tmem_base, dah, dbl, idesc, 1u);
              }
              ptx::mma_commit_elect(&w_empty[buf]);
            }
            ptx::mma_commit_elect(acc_full);
            CH_T2(grp, oi, 0);
          }
          wi += o.nchunks;
          ptx::mbar_wait(&p_full[pi & 1], (uint32_t)((pi >> 1) & 1));
          const float* par = reinterpret_cast<const float*>(s_par + (pi & 1) * kChainParSlot);
          const float* bias = par + 10 * o.par_dw_c;
          const float* alpha = bias + Np;
          ptx::mbar_wait(acc_full, (uint32_t)acc_phase);
          acc_phase ^= 1;
          ptx::tc_fence_after_sync();
          CH_T2(grp, oi, 1);
          // epilogue: warp w owns TMEM lanes 32 (w & 3) .. + 31 (rows) and the column group w >> 2
          const int row = (warp & 3) * 32 + lane, cg = warp >> 2;
          const int cw = ((Np >> 2) + 15) / 16 * 16;
          const int rows_out = G * o.out.H * o.out.W;
          const int c_end = min(Np, (cg + 1) * cw);
          const int skf = o.has_skip ? (o.skip.fmt == CH_P16 ? 2 : 1) : 0;
          for (int c0 = cg * cw; c0 < c_end; c0 += 16) {
            float v[16];
            ptx::tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
            if (row < rows_out) {
              if (o.out.fmt == CH_P16) {
                if (skf == 2) epilogue16<true, 2>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
                else if (skf == 1) epilogue16<true, 1>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
                else epilogue16<true, 0>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
              } else {
                if (skf == 2) epilogue16<false, 2>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
                else if (skf == 1) epilogue16<false, 1>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
                else epilogue16<false, 0>(arena, v, row, c0, bias, alpha, o.act, o.out, o.skip, o.skip_c);
              }
              if (o.out2.off >= 0) {                             // a second, F32 copy (the planner never overwrites the residual in place then)
                if (skf == 2) epilogue16<false, 2>(arena, v, row, c0, bias, alpha, o.act, o.out2, o.skip, o.skip_c);
                else if (skf == 1) epilogue16<false, 1>(arena, v, row, c0, bias, alpha, o.act, o.out2, o.skip, o.skip_c);
                else epilogue16<false, 0>(arena, v, row, c0, bias, alpha, o.act, o.out2, o.skip, o.skip_c);
              }
            }
          }
        }
        CH_T2(grp, oi, 2);
        if (o.no_barrier) { CH_T(grp, 1 + oi); continue; }       // (a POOL next to the depthwise of the same step: disjoint data)
        // ---- end of the phase: generic-proxy writes -> visible to the tensor core / later phases ----
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before_sync();
        workers_sync();
        ptx::tc_fence_after_sync();
        CH_T(grp, 1 + oi);
        if (o.par_release) {
          if (tid == 0) mbar_arrive(&p_empty[pi & 1]);
          ++pi;
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 128);
}

}  // namespace

cudaError_t chain_init() { return cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem); }

bool chain_enabled() {
  static const bool on = [] { const char* e = getenv("FDL_CHAIN"); return e ? atoi(e) != 0 : true; }();
  return on;
}

cudaError_t launch_chain(const ChainArgs& a, cudaStream_t stream) {
  const int groups = (a.B + a.items - 1) / a.items;
  int grid = persist_sms();
  if (grid > groups) grid = groups;
  cudaError_t e = launch_pdl(chain_kernel, dim3(grid), dim3(kChainThreads + 32), (size_t)kChainSmem, stream, a);
  count_launch();
  return e;
}

}  // namespace fdl
