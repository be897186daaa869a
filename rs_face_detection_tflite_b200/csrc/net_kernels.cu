// net_kernels.cu -- network kernels, fp32 path.  See net_kernels.cuh.
//
// fused_conv_kernel is an implicit GEMM  out[M = B*OH*OW][N = Cout] = A[M][K] * W[K][N]  whose A tile
// is produced on chip:
//   mode 0 (CONV_2D, any kh/kw/stride, SAME/VALID): im2col gather, K = kh*kw*Cin;
//   mode 1 (BlazeBlock):  A = DEPTHWISE_CONV_2D 3x3 (+bias) of the input, K = Cin, so the depthwise
//           result never touches HBM;
// and whose epilogue applies bias, the residual branch (identity / MAX_POOL 2x2 / zero channel PAD)
// and RELU / PRELU before the single store of the block output.
#include "net_kernels.cuh"

#include <atomic>

#include "plan.h"

namespace fdl {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count_value() { return g_launches.load(std::memory_order_relaxed); }

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float apply_act(float v, int act, const float* alpha, int n) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_PRELU) return v >= 0.f ? v : v * __ldg(alpha + n);
  return v;
}

template <int BM>
__global__ void __launch_bounds__(kThreads) fused_conv_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int OHW = a.out.H * a.out.W;
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const long long M = (long long)nb * OHW;
  const long long m0 = (long long)blockIdx.x * BM;
  if (m0 >= M) return;
  const int K = a.K, K4 = a.K4, ldA = K4 + 4;
  const int Cin = a.in.C, IH = a.in.H, IW = a.in.W;

  // ---------------- phase A: build the [BM][K4] operand tile ----------------
  if (a.mode == 0) {
    const int kwc = a.kw * Cin;
    for (int idx = tid; idx < BM * K4; idx += kThreads) {
      int px = idx / K4, k = idx - px * K4;
      float v = 0.f;
      long long m = m0 + px;
      if (k < K && m < M) {
        int b = (int)(m / OHW);
        int pix = (int)(m - (long long)b * OHW);
        int oy = pix / a.out.W, ox = pix - oy * a.out.W;
        int ky = k / kwc, r = k - ky * kwc;
        int kx = r / Cin, ci = r - kx * Cin;
        int iy = oy * a.stride - a.pad_t + ky, ix = ox * a.stride - a.pad_l + kx;
        if (iy >= 0 && iy < IH && ix >= 0 && ix < IW)
          v = __ldg(a.in.p + (long long)b * a.in.bstride + ((long long)iy * IW + ix) * Cin + ci);
      }
      smem[px * ldA + k] = v;
    }
  } else {
    for (int idx = tid; idx < BM * K4; idx += kThreads) {
      int px = idx / K4, ci = idx - px * K4;
      float v = 0.f;
      long long m = m0 + px;
      if (ci < K && m < M) {
        int b = (int)(m / OHW);
        int pix = (int)(m - (long long)b * OHW);
        int oy = pix / a.out.W, ox = pix - oy * a.out.W;
        const float* src = a.in.p + (long long)b * a.in.bstride + ci;
        int iy0 = oy * a.stride - a.pad_t, ix0 = ox * a.stride - a.pad_l;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          int iy = iy0 + ky;
          if (iy < 0 || iy >= IH) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            int ix = ix0 + kx;
            if (ix < 0 || ix >= IW) continue;
            acc = fmaf(__ldg(src + ((long long)iy * IW + ix) * Cin), __ldg(a.w_dw + (ky * 3 + kx) * Cin + ci), acc);
          }
        }
        v = acc + __ldg(a.b_dw + ci);
      }
      smem[px * ldA + ci] = v;
    }
  }
  __syncthreads();

  // ---------------- phase B: GEMM, 4x4 register tile per thread ----------------
  constexpr int TY = BM / 4;             // thread rows
  constexpr int TX = kThreads / TY;      // thread cols; each covers 4 output channels per pass
  constexpr int BNp = TX * 4;            // channels per pass
  const int tx = tid % TX, ty = tid / TX;
  const int N = a.N, Npad = a.Npad;
  for (int n0 = blockIdx.y * BNp; n0 < Npad; n0 += gridDim.y * BNp) {
    const int n = n0 + tx * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (n < Npad) {
      const float* wp = a.w + n;
      for (int k = 0; k < K4; k += 4) {
        float4 av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(&smem[(ty * 4 + i) * ldA + k]);
        float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(k + 0) * Npad));
        float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(k + 1) * Npad));
        float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(k + 2) * Npad));
        float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + (long long)(k + 3) * Npad));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][0] = fmaf(av[i].x, w0.x, acc[i][0]); acc[i][1] = fmaf(av[i].x, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av[i].x, w0.z, acc[i][2]); acc[i][3] = fmaf(av[i].x, w0.w, acc[i][3]);
          acc[i][0] = fmaf(av[i].y, w1.x, acc[i][0]); acc[i][1] = fmaf(av[i].y, w1.y, acc[i][1]);
          acc[i][2] = fmaf(av[i].y, w1.z, acc[i][2]); acc[i][3] = fmaf(av[i].y, w1.w, acc[i][3]);
          acc[i][0] = fmaf(av[i].z, w2.x, acc[i][0]); acc[i][1] = fmaf(av[i].z, w2.y, acc[i][1]);
          acc[i][2] = fmaf(av[i].z, w2.z, acc[i][2]); acc[i][3] = fmaf(av[i].z, w2.w, acc[i][3]);
          acc[i][0] = fmaf(av[i].w, w3.x, acc[i][0]); acc[i][1] = fmaf(av[i].w, w3.y, acc[i][1]);
          acc[i][2] = fmaf(av[i].w, w3.z, acc[i][2]); acc[i][3] = fmaf(av[i].w, w3.w, acc[i][3]);
        }
      }
      // ---------------- epilogue ----------------
      float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + n));
      const float bias4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int b = (int)(m / OHW);
        int pix = (int)(m - (long long)b * OHW);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias4[j];
        if (a.has_skip) {
          const float* sp = a.skip.p + (long long)b * a.skip.bstride;
          if (a.skip_pool) {
            int oy = pix / a.out.W, ox = pix - oy * a.out.W;
            const float* s00 = sp + ((long long)(2 * oy) * a.skip.W + 2 * ox) * a.skip.C;
            const float* s10 = s00 + (long long)a.skip.W * a.skip.C;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              int c = n + j;
              if (c < a.skip_c) {
                float mx = fmaxf(fmaxf(__ldg(s00 + c), __ldg(s00 + a.skip.C + c)), fmaxf(__ldg(s10 + c), __ldg(s10 + a.skip.C + c)));
                v[j] += mx;
              }
            }
          } else {
            const float* s0 = sp + (long long)pix * a.skip.C;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < a.skip_c) v[j] += __ldg(s0 + n + j);
          }
        }
        float* op = a.out.p + (long long)b * a.out.bstride + (long long)pix * N + n;
        if ((N & 3) == 0) {
          float4 o;
          o.x = apply_act(v[0], a.act, a.alpha, n + 0); o.y = apply_act(v[1], a.act, a.alpha, n + 1);
          o.z = apply_act(v[2], a.act, a.alpha, n + 2); o.w = apply_act(v[3], a.act, a.alpha, n + 3);
          *reinterpret_cast<float4*>(op) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (n + j < N) op[j] = apply_act(v[j], a.act, a.alpha, n + j);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone ops (only reached for graphs whose structure the fusion patterns do not cover; the
// five dense graphs use RESIZE(+ADD) only).
__global__ void elementwise_kernel(const EltArgs a) {
  const int C = a.out.C, OW = a.out.W, OH = a.out.H;
  int nb = a.B;
  if (a.n_active) nb = min(nb, *a.n_active);
  const long long per = (long long)OH * OW * C;
  const long long total = per * nb;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int b = (int)(idx / per);
    long long r = idx - (long long)b * per;
    int c = (int)(r % C);
    int pix = (int)(r / C);
    int oy = pix / OW, ox = pix - oy * OW;
    const float* ip = a.in.p + (long long)b * a.in.bstride;
    float v = 0.f;
    switch (a.kind) {
      case STEP_DW: {
        int iy0 = oy * a.stride - a.pad_t, ix0 = ox * a.stride - a.pad_l;
        float acc = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
          int iy = iy0 + ky;
          if (iy < 0 || iy >= a.in.H) continue;
          for (int kx = 0; kx < 3; ++kx) {
            int ix = ix0 + kx;
            if (ix < 0 || ix >= a.in.W) continue;
            acc = fmaf(__ldg(ip + ((long long)iy * a.in.W + ix) * C + c), __ldg(a.w_dw + (ky * 3 + kx) * C + c), acc);
          }
        }
        v = acc + __ldg(a.b_dw + c);
        break;
      }
      case STEP_POOL: {
        const float* s = ip + ((long long)(2 * oy) * a.in.W + 2 * ox) * C + c;
        v = fmaxf(fmaxf(__ldg(s), __ldg(s + C)), fmaxf(__ldg(s + (long long)a.in.W * C), __ldg(s + (long long)a.in.W * C + C)));
        break;
      }
      case STEP_PADC:
        v = c < a.in.C ? __ldg(ip + (long long)pix * a.in.C + c) : 0.f;
        break;
      case STEP_ADD:
        v = __ldg(ip + (long long)pix * C + c) + __ldg(a.other.p + (long long)b * a.other.bstride + (long long)pix * C + c);
        break;
      case STEP_ACT:
        v = __ldg(ip + (long long)pix * C + c);
        break;
      case STEP_D2S: {
        // DEPTH_TO_SPACE, block size a.stride: out[oy, ox, c] = in[oy / bs, ox / bs, ((oy % bs) * bs + ox % bs) * C + c]
        const int bs = a.stride;
        const int iy = oy / bs, ix = ox / bs;
        v = __ldg(ip + ((long long)iy * a.in.W + ix) * a.in.C + ((oy - iy * bs) * bs + (ox - ix * bs)) * C + c);
        break;
      }
      case STEP_RESIZE: {
        // TFLite RESIZE_BILINEAR, align_corners = false, half_pixel_centers = true
        float sy = (oy + 0.5f) * ((float)a.in.H / (float)OH) - 0.5f;
        float sx = (ox + 0.5f) * ((float)a.in.W / (float)OW) - 0.5f;
        float fy = floorf(sy), fx = floorf(sx);
        int y0 = max((int)fy, 0), y1 = min((int)ceilf(sy), a.in.H - 1);
        int x0 = max((int)fx, 0), x1 = min((int)ceilf(sx), a.in.W - 1);
        float wy = sy - fy, wx = sx - fx;
        float v00 = __ldg(ip + ((long long)y0 * a.in.W + x0) * C + c), v01 = __ldg(ip + ((long long)y0 * a.in.W + x1) * C + c);
        float v10 = __ldg(ip + ((long long)y1 * a.in.W + x0) * C + c), v11 = __ldg(ip + ((long long)y1 * a.in.W + x1) * C + c);
        v = (v00 * (1.f - wx) + v01 * wx) * (1.f - wy) + (v10 * (1.f - wx) + v11 * wx) * wy;
        if (a.has_other) v += __ldg(a.other.p + (long long)b * a.other.bstride + (long long)pix * C + c);
        break;
      }
      default: break;
    }
    v = apply_act(v, a.act, a.alpha, c);
    a.out.p[(long long)b * a.out.bstride + (long long)pix * C + c] = v;
  }
}

constexpr int kMaxSmem = 200 * 1024;

}  // namespace

cudaError_t net_kernels_init() {
  cudaError_t e = cudaFuncSetAttribute(fused_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fused_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  if (e != cudaSuccess) return e;
  e = stem_kernels_init();
  if (e != cudaSuccess) return e;
  e = stem_tc_init();
  if (e != cudaSuccess) return e;
  e = pw_tc_init();
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fused_conv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
}

cudaError_t launch_fused_conv(const ConvArgs& a, cudaStream_t stream) {
  const long long M = (long long)a.B * a.out.H * a.out.W;
  if (M <= 0) return cudaSuccess;
  if (stem_tc_supported(a)) return launch_stem_tc(a, stream);
  if (stem_supported(a)) return launch_stem_conv(a, stream);
  const int ldA = a.K4 + 4;
  // pick the pixel tile: as large as shared memory allows, smaller when the problem is tiny
  int BM = 64;
  while (BM > 16 && ((size_t)BM * ldA * 4 > (size_t)kMaxSmem || M <= BM * 74)) BM >>= 1;
  size_t smem = (size_t)BM * ldA * sizeof(float);
  if (smem > (size_t)kMaxSmem) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((M + BM - 1) / BM), 1, 1);
  // few pixel tiles but many output channels (dense tails: M = batch): split the channel passes across CTAs
  const int BNp = (kThreads / (BM / 4)) * 4;
  if (grid.x < 148) grid.y = (unsigned)((a.Npad + BNp - 1) / BNp);
  if (BM == 64) fused_conv_kernel<64><<<grid, kThreads, smem, stream>>>(a);
  else if (BM == 32) fused_conv_kernel<32><<<grid, kThreads, smem, stream>>>(a);
  else fused_conv_kernel<16><<<grid, kThreads, smem, stream>>>(a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_elementwise(const EltArgs& a, cudaStream_t stream) {
  const long long total = (long long)a.B * a.out.H * a.out.W * a.out.C;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  elementwise_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  count_launch();
  return cudaGetLastError();
}

}  // namespace fdl
