// conv_tc_kernel.cu -- general tensor-core convolution (sm_100a: tcgen05 + TMEM), for everything the TMA-tiled
// BlazeBlock kernel does not take: CONV_2D 1x1 / 2x2 s2 / dense tails (k x k VALID over the whole map) / heads, and
// BlazeBlocks on small feature maps (16x16 .. 2x2) where the parallelism has to come from the batch.
//
//   out[M = B*OH*OW pixels][N] = act( A[M][K] * W[K][N] + bias + skip )
//
// Pixels are flattened over the batch, so a 128-row MMA tile spans as many images as it needs.  The A tile is
// gathered by the CTA's threads with 16-byte loads (im2col on the fly, or the depthwise 3x3 of a BlazeBlock
// evaluated in registers), split into tf32 hi / lo parts (see mma_kernels.cu) and stored in shared memory in the
// UMMA K-major core-matrix layout, 32 K-values at a time, double buffered: while tcgen05.mma consumes chunk c the
// threads gather chunk c+1.  The accumulator lives in TMEM; the epilogue (tcgen05.ld, + bias, + residual,
// RELU / PRELU) writes straight to global memory, which also covers the aliased head outputs ([B,N,16] / [B,N,1]).
#include <cuda_runtime.h>

#include <cstdlib>

#include "mma_kernels.cuh"
#include "pdl.h"
#include "plan.h"
#include "sm100_ptx.cuh"

namespace fdl {

void count_launch();

namespace {

constexpr int kThreads = 256;
// K values per chunk: KC = 4 * QC, QC = channel quads (A planes) per chunk.  Two instantiations: QC = 8 (32 K values per
// chunk) and QC = 4 (16 per chunk: half the shared memory per CTA, so twice the CTAs per SM -- these CTAs are single
// latency chains (gather -> planes -> MMA -> epilogue), and the SM hides the chain of one behind the others).
constexpr int kPlaneBytes = 128 * 16 + 16;           // one A plane: 128 pixels x 16 B (+16 B bank skew)

struct Layout { int bias, alpha, dw, a0, w0, a_stage, w_stage, total; };

__host__ __device__ inline int align_up_c(int v, int a) { return (v + a - 1) / a * a; }

__host__ __device__ inline Layout layout(int Nt, int wsplit, int dw_c, int QC) {
  const int kABytes = QC * kPlaneBytes;              // hi (or lo) planes of one chunk
  Layout L;
  int off = 64;
  L.bias = off; off += Nt * 4;
  L.alpha = off; off += Nt * 4;
  off = align_up_c(off, 16);
  L.dw = off; off += dw_c * 10 * 4;                 // depthwise weights [9][C] + bias [C] (BLOCK mode)
  off = align_up_c(off, 128);
  L.a_stage = 2 * kABytes;                          // hi + lo
  L.a0 = off; off += 2 * L.a_stage;
  off = align_up_c(off, 128);
  L.w_stage = wsplit * QC * Nt * 16;
  L.w0 = off; off += 2 * L.w_stage;
  L.total = align_up_c(off, 128);
  return L;
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
  a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}

template <int QC>
__global__ void __launch_bounds__(kThreads, QC == 4 ? 3 : 0) conv_tc_kernel(const ConvTcArgs a) {
  constexpr int KC = 4 * QC;                         // K values per chunk
  constexpr int kABytes = QC * kPlaneBytes;          // hi (or lo) planes of one chunk
  constexpr int PPT = QC / 2;                        // pixels per thread and chunk: 256 threads = (256 / QC) pixel slots x QC quads
  constexpr int PSTEP = kThreads / QC;               // pixel slots
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int Nt = a.Nt;
  const Layout L = layout(Nt, a.wsplit, a.mode == 1 ? a.in.C : 0, QC);
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem);           // [2]: chunk buffer free / all done
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + 16);        // [2]: weight chunk landed (bulk copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 48);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  float* s_alpha = reinterpret_cast<float*>(smem + L.alpha);
  float* s_dw = reinterpret_cast<float*>(smem + L.dw);

  const int OHW = a.out.H * a.out.W;
  const long long m0 = (long long)blockIdx.x * 128;
  const int nt = blockIdx.y;                       // N tile
  const int n_base = nt * Nt;

  // ---- prologue: nothing here depends on the previous launch (PDL, see pdl.h) ----
  if (tid == 0) {
    ptx::mbar_init(&mma_bar[0], 1);
    ptx::mbar_init(&mma_bar[1], 1);
    ptx::mbar_init(&w_bar[0], 1);
    ptx::mbar_init(&w_bar[1], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  for (int i = tid; i < Nt; i += kThreads) {
    const int n = n_base + i;
    s_bias[i] = n < a.N ? a.bias[n] : 0.f;
    s_alpha[i] = (a.alpha && n < a.N) ? a.alpha[n] : 0.f;
  }
  if (a.mode == 1) {
    const int C = a.in.C;
    for (int i = tid; i < 9 * C; i += kThreads) s_dw[i] = a.w_dw[i];
    for (int i = tid; i < C; i += kThreads) s_dw[9 * C + i] = a.b_dw[i];
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  pdl_launch_dependents();
  pdl_wait();                                      // the previous launch's activations (and *n_active) are visible from here on
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const long long M = (long long)nb * OHW;
  if (m0 < M) {                                    // CTA-uniform

  // this thread's gather slots: quad j (fixed) of pixels p0 + PSTEP*i
  const int j = tid & (QC - 1);
  const int p0 = tid / QC;
  int pb[PPT], py[PPT], px[PPT];
  bool pv[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    long long m = m0 + p0 + PSTEP * i;
    pv[i] = m < M;
    long long mm = pv[i] ? m : 0;
    pb[i] = (int)(mm / OHW);
    int pix = (int)(mm - (long long)pb[i] * OHW);
    py[i] = pix / a.out.W;
    px[i] = pix - py[i] * a.out.W;
  }
  const int Cin = a.in.C, IH = a.in.H, IW = a.in.W;
  const int nchunks = a.Kp / KC;
  const uint32_t idesc = ptx::umma_idesc_tf32(128, Nt);
  const uint32_t w_lbo = (uint32_t)Nt * 16u;
  const float4* wsrc = reinterpret_cast<const float4*>(a.w_tc) + (size_t)nt * a.wsplit * (a.Kp / 4) * Nt;   // this N tile
  const int w4_per_chunk = QC * Nt;                 // float4s of one (hi or lo) chunk
  const size_t w4_lo_off = (size_t)(a.Kp / 4) * Nt; // offset of the lo copy inside the N tile

  // im2col gather of one chunk into registers: all four 16-byte loads are issued back to back (branch-free), so
  // their latencies overlap; the loop below prefetches chunk c+1 while chunk c is converted, stored and multiplied.
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gather_conv = [&](int c, float4 (&v)[PPT]) {
    const int k = c * KC + 4 * j;
    const int kwc = a.kw * Cin;
    const int ky = k / kwc, r = k - ky * kwc;
    const int kx = r / Cin, ci = r - kx * Cin;
    const bool kok = k < a.K;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int iy = py[i] * a.stride - a.pad_t + ky, ix = px[i] * a.stride - a.pad_l + kx;
      const bool ok = kok && pv[i] && iy >= 0 && iy < IH && ix >= 0 && ix < IW;
      const float4* ptr = reinterpret_cast<const float4*>(a.in.p + (long long)pb[i] * a.in.bstride + ((long long)iy * IW + ix) * Cin + ci);
      v[i] = ok ? __ldg(ptr) : zero4;
    }
  };
  // same for input channel counts that are not a multiple of 4 (the RGB stems): element-wise gather, 16 scalar loads
  auto gather_scalar = [&](int c, float4 (&v)[PPT]) {
    const int k0 = c * KC + 4 * j;
    const int kwc = a.kw * Cin;
    int dy[4], dx[4], dc[4];
    bool kok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      const int ky = k / kwc, r = k - ky * kwc;
      dy[e] = ky; dx[e] = r / Cin; dc[e] = r - dx[e] * Cin;
      kok[e] = k < a.K;
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float* img = a.in.p + (long long)pb[i] * a.in.bstride;
      const int iy0 = py[i] * a.stride - a.pad_t, ix0 = px[i] * a.stride - a.pad_l;
      float t[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int iy = iy0 + dy[e], ix = ix0 + dx[e];
        const bool ok = kok[e] && pv[i] && iy >= 0 && iy < IH && ix >= 0 && ix < IW;
        t[e] = ok ? __ldg(img + ((long long)iy * IW + ix) * Cin + dc[e]) : 0.f;
      }
      v[i] = make_float4(t[0], t[1], t[2], t[3]);
    }
  };
  float4 pre[PPT];
  if (a.mode == 0) gather_conv(0, pre);
  else if (a.mode == 2) gather_scalar(0, pre);

  // Weights of a chunk: already in UMMA order in global memory -> one or two bulk async copies, requested ONE CHUNK AHEAD
  // (a CTA is a single latency chain; a load requested in the iteration that needs it costs a full L2 round trip per chunk).
  auto issue_w = [&](int c) {
    const int wb = c & 1;
    float4* dst = reinterpret_cast<float4*>(smem + L.w0 + wb * L.w_stage);
    const uint32_t bytes = (uint32_t)w4_per_chunk * 16u;
    ptx::mbar_arrive_expect_tx(&w_bar[wb], bytes * (uint32_t)a.wsplit);
    ptx::bulk_load_1d(dst, wsrc + (size_t)c * w4_per_chunk, bytes, &w_bar[wb]);
    if (a.wsplit == 2) ptx::bulk_load_1d(dst + w4_per_chunk, wsrc + w4_lo_off + (size_t)c * w4_per_chunk, bytes, &w_bar[wb]);
  };
  if (tid == 0) issue_w(0);

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    // the MMAs that read this buffer two chunks ago must have completed
    if (c >= 2) ptx::mbar_wait(&mma_bar[buf], (uint32_t)(((c >> 1) - 1) & 1));
    uint8_t* s_a = smem + L.a0 + buf * L.a_stage;
    float4* s_w = reinterpret_cast<float4*>(smem + L.w0 + buf * L.w_stage);
    if (tid == 0 && c + 1 < nchunks) {
      // the other weight buffer was last read by the MMAs of chunk c-1 (issued a moment ago by this thread)
      if (c >= 1) ptx::mbar_wait(&mma_bar[buf ^ 1], (uint32_t)(((c - 1) >> 1) & 1));
      issue_w(c + 1);
    }
    // ---- the A chunk ----
    const int k = c * KC + 4 * j;                   // first K index of this thread's quad
    float4 cur[PPT];
    if (a.mode != 1) {
#pragma unroll
      for (int i = 0; i < PPT; ++i) cur[i] = pre[i];
      if (c + 1 < nchunks) {
        if (a.mode == 0) gather_conv(c + 1, pre);
        else gather_scalar(c + 1, pre);
      }
    } else {
      // depthwise 3x3 (+bias) of channel quad k..k+3 at each of the four pixels; loads predicated, not branched
      const bool kok = k < a.K;
      const float4 bdw = kok ? *reinterpret_cast<const float4*>(s_dw + 9 * Cin + k) : zero4;
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        const float* img = a.in.p + (long long)pb[i] * a.in.bstride + k;
        const int iy0 = py[i] * a.stride - a.pad_t, ix0 = px[i] * a.stride - a.pad_l;
        float4 t[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int iy = iy0 + ky, ix = ix0 + kx;
            const bool ok = kok && pv[i] && iy >= 0 && iy < IH && ix >= 0 && ix < IW;
            t[ky * 3 + kx] = ok ? __ldg(reinterpret_cast<const float4*>(img + ((long long)iy * IW + ix) * Cin)) : zero4;
          }
        float4 v = (pv[i] && kok) ? bdw : zero4;
        if (kok) {
#pragma unroll
          for (int q = 0; q < 9; ++q) fma4(v, t[q], *reinterpret_cast<const float4*>(s_dw + q * Cin + k));
        }
        cur[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float4 v = cur[i];
      float4 hi, lo;
      hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
      lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
      const int p = p0 + PSTEP * i;
      *reinterpret_cast<float4*>(s_a + j * kPlaneBytes + p * 16) = hi;
      *reinterpret_cast<float4*>(s_a + kABytes + j * kPlaneBytes + p * 16) = lo;
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp_u == 0) {                               // warp 0, converged: one elected lane issues (operands stay uniform, see mma_f16_elect)
      ptx::mbar_wait(&w_bar[buf], (uint32_t)((c >> 1) & 1));
      ptx::tc_fence_after_sync();
      const uint32_t ahi = ptx::smem_u32(s_a), alo = ahi + kABytes, wb = ptx::smem_u32(s_w);
      for (int pass = 0; pass < (a.wsplit == 2 ? 3 : 2); ++pass) {
        const uint32_t a_base = pass == 0 ? alo : ahi;
        const uint32_t b_base = pass == 2 ? wb + (uint32_t)(w4_per_chunk * 16) : wb;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          uint64_t ad = ptx::umma_desc_kmajor(a_base + (uint32_t)(ks * 2 * kPlaneBytes), kPlaneBytes, 128);
          uint64_t bd = ptx::umma_desc_kmajor(b_base + (uint32_t)ks * 2u * w_lbo, w_lbo, 128);
          ptx::mma_tf32_elect(tmem_base, ad, bd, idesc, (c > 0 || pass > 0 || ks > 0) ? 1u : 0u);
        }
      }
      ptx::mma_commit_elect(&mma_bar[buf]);         // frees this buffer; the last commit also signals the epilogue
    }
  }

  // ---- epilogue ----
  if (warp < 4) {
    const int last = nchunks - 1;
    ptx::mbar_wait(&mma_bar[last & 1], (uint32_t)((last >> 1) & 1));
    ptx::tc_fence_after_sync();
    const int p = tid;
    const long long m = m0 + p;
    const bool valid = m < M;
    const long long mm = valid ? m : 0;
    const int b = (int)(mm / OHW);
    const int pix = (int)(mm - (long long)b * OHW);
    const int N = a.N;
    float* op = a.out.p + (long long)b * a.out.bstride + (long long)pix * N;
    const float* sp = nullptr;
    long long srow = 0;
    if (a.has_skip && valid) {
      if (a.skip_pool) {
        const int oy = pix / a.out.W, ox = pix - oy * a.out.W;
        sp = a.skip.p + (long long)b * a.skip.bstride + ((long long)(2 * oy) * a.skip.W + 2 * ox) * a.skip.C;
        srow = (long long)a.skip.W * a.skip.C;
      } else {
        sp = a.skip.p + (long long)b * a.skip.bstride + (long long)pix * a.skip.C;
      }
    }
    // Residual values are fetched 64 channels at a time with 16-byte loads, ALL issued before the accumulator columns
    // are read, so the batch costs one global-memory round trip instead of one per channel quad (this kernel runs the
    // small feature maps, where a CTA is a single latency chain).
    const bool vec_skip = sp && (a.skip.C & 3) == 0 && (a.skip_c & 3) == 0 && (N & 3) == 0;
    constexpr int EB = QC == 4 ? 32 : 64;          // residual channels fetched per batch (register budget of the variant)
    for (int b0 = 0; b0 < Nt; b0 += EB) {
      float4 sk[EB / 4];
#pragma unroll
      for (int j = 0; j < EB / 4; ++j) {
        sk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ch = n_base + b0 + 4 * j;
        if (vec_skip && b0 + 4 * j < Nt && ch < a.skip_c) {
          if (a.skip_pool) {
            const float4 s00 = __ldg(reinterpret_cast<const float4*>(sp + ch)), s01 = __ldg(reinterpret_cast<const float4*>(sp + a.skip.C + ch));
            const float4 s10 = __ldg(reinterpret_cast<const float4*>(sp + srow + ch)), s11 = __ldg(reinterpret_cast<const float4*>(sp + srow + a.skip.C + ch));
            sk[j] = make_float4(fmaxf(fmaxf(s00.x, s01.x), fmaxf(s10.x, s11.x)), fmaxf(fmaxf(s00.y, s01.y), fmaxf(s10.y, s11.y)),
                                fmaxf(fmaxf(s00.z, s01.z), fmaxf(s10.z, s11.z)), fmaxf(fmaxf(s00.w, s01.w), fmaxf(s10.w, s11.w)));
          } else {
            sk[j] = __ldg(reinterpret_cast<const float4*>(sp + ch));
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < EB; cc += 16) {
        const int c0 = b0 + cc;
        if (c0 >= Nt) break;
        float v[16];
        ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (!valid) continue;
#pragma unroll
        for (int q4 = 0; q4 < 16; q4 += 4) {
          const int n = n_base + c0 + q4;
          if (n >= N) break;
          float o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o4[e] = v[q4 + e] + s_bias[c0 + q4 + e];
          if (vec_skip) {
            const float4 s4 = sk[(cc + q4) >> 2];
            o4[0] += s4.x; o4[1] += s4.y; o4[2] += s4.z; o4[3] += s4.w;
          } else if (sp) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ch = n + e;
              if (ch < a.skip_c) {
                if (a.skip_pool) o4[e] += fmaxf(fmaxf(__ldg(sp + ch), __ldg(sp + a.skip.C + ch)), fmaxf(__ldg(sp + srow + ch), __ldg(sp + srow + a.skip.C + ch)));
                else o4[e] += __ldg(sp + ch);
              }
            }
          }
          if (a.act == ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] = fmaxf(o4[e], 0.f);
          } else if (a.act == ACT_PRELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] = o4[e] >= 0.f ? o4[e] : o4[e] * s_alpha[c0 + q4 + e];
          }
          if ((N & 3) == 0) {
            *reinterpret_cast<float4*>(op + n) = make_float4(o4[0], o4[1], o4[2], o4[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (n + e < N) op[n + e] = o4[e];
          }
        }
      }
    }
  }
  }  // m0 < M
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

}  // namespace

cudaError_t conv_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

bool conv_tc_supported(const Step& s) {
  if (s.kind != STEP_CONV && s.kind != STEP_BLOCK) return false;
  if (s.w_tc < 0) return false;
  if (s.in.C % 4 != 0 && s.kind != STEP_CONV) return false;   // odd channel counts: scalar im2col gather (mode 2), CONV only
  if (s.in.C == 3) return false;   // RGB stems: the FFMA stem kernel (stem_kernel.cu) is faster than a scalar gather (measured)
  if (s.in.offset != 0) return false;
  if (s.kind == STEP_BLOCK && (s.in.C > 256)) return false;
  if (s.skip.tensor >= 0 && s.skip.offset != 0) return false;
  Layout L = layout(s.Nt, s.wsplit, s.kind == STEP_BLOCK ? s.in.C : 0, 4);
  return L.total <= 200 * 1024;
}

cudaError_t launch_conv_tc(const ConvTcArgs& a0, cudaStream_t stream) {
  ConvTcArgs a = a0;
  a.tmem_cols = a.Nt <= 32 ? 32 : (a.Nt <= 64 ? 64 : 128);
  const long long M = (long long)a.B * a.out.H * a.out.W;
  if (M <= 0) return cudaSuccess;
  const int dw_c = a.mode == 1 ? a.in.C : 0;
  const Layout L8 = layout(a.Nt, a.wsplit, dw_c, 8), L4 = layout(a.Nt, a.wsplit, dw_c, 4);
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)a.n_tiles, 1);
  // Chunk size: 32 K-values per chunk unless that leaves CTAs waiting for a slot -- then 16 per chunk (half the shared
  // memory, twice the resident CTAs).  FDL_CONV_QC = 4 / 8 forces one (A/B timing).
  static const int qc_env = getenv("FDL_CONV_QC") ? atoi(getenv("FDL_CONV_QC")) : 0;
  // resident CTAs per SM: shared memory, and registers (128 per thread -> 2 CTAs for <8>; 80 -> 3 for <4>, see __launch_bounds__)
  auto per_sm = [](int smem, int reg_cap) { int n = (228 * 1024) / (smem + 1024); return n > reg_cap ? reg_cap : n; };
  const long long ctas = (long long)grid.x * grid.y;
  bool small = L8.total > 200 * 1024 || (ctas > 148LL * per_sm(L8.total, 2) && per_sm(L4.total, 3) > per_sm(L8.total, 2));
  if (qc_env == 8 && L8.total <= 200 * 1024) small = false;
  if (qc_env == 4) small = true;
  cudaError_t e = small ? launch_pdl(conv_tc_kernel<4>, grid, dim3(kThreads), (size_t)L4.total, stream, a)
                        : launch_pdl(conv_tc_kernel<8>, grid, dim3(kThreads), (size_t)L8.total, stream, a);
  count_launch();
  return e;
}

}  // namespace fdl
