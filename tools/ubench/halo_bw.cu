// halo_bw.cu -- what the memory system gives the BlazeBlock kernel's access patterns (B200, sm_100a).
// Tensor [B][H][W][C] f32 NHWC (default 256 x 128 x 128 x 24 = 403 MB, the detector's dominant stage).  Data movement only:
//   0  ldg      plain 16-byte grid-stride copy (read + write)                         -- the STREAM-style reference
//   1  tile     4-D TMA loads of 10 x 18 x C halo tiles, NS stages per CTA, 2 CTAs/SM  -- block_ws_kernel's load path, loads only
//   2  tile+st  the same + a 4-D TMA store of every 8 x 16 x C tile                    -- its whole data-movement skeleton
//   3  rows     1-D bulk loads of whole image rows (W*C*4 contiguous bytes), NS deep    -- row-streaming, loads only
//   4  rows+st  the same + a 1-D bulk store of every row                               -- row-streaming skeleton (read + write)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rs_face_detection_tflite_b200/csrc halo_bw.cu -o halo_bw
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "sm100_ptx.cuh"

using namespace fdl;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void ldg_copy(const float4* __restrict__ in, float4* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

constexpr int TH = 8, TW = 16, ITH = 10, ITW = 18;

__global__ void __launch_bounds__(64) tile_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int B, int H, int W, int C,
                                                  int NS, int do_store, int CPin) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  const int stage_bytes = (ITH * ITW * CPin * 4 + 127) / 128 * 128;
  uint8_t* buf = smem + 1024;
  const int tx_n = W / TW, ty_n = H / TH, per_img = tx_n * ty_n, ntiles = B * per_img;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_mbar_init();
    const int my = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto issue = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x, s = it % NS;
      const int b = tile / per_img, r = tile % per_img, ty = r / tx_n, tx = r % tx_n;
      ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(ITH * ITW * CPin * 4));
      ptx::tma_load_4d(buf + s * stage_bytes, &tm_in, &full[s], 0, tx * TW - 1, ty * TH - 1, b);
    };
    for (int it = 0; it < NS && it < my; ++it) issue(it);
    for (int it = 0; it < my; ++it) {
      const int s = it % NS;
      ptx::mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
      if (do_store) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int b = tile / per_img, r = tile % per_img, ty = r / tx_n, tx = r % tx_n;
        // store the first 8 x 16 pixels' worth of the stage (contents do not matter here)
        ptx::tma_store_4d(&tm_out, buf + s * stage_bytes, 0, tx * TW, ty * TH, b);
        ptx::tma_store_commit();
        ptx::tma_store_wait_read0();
      }
      if (it + NS < my) issue(it + NS);
    }
    if (do_store) ptx::tma_store_wait_all0();
  }
}

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(ptx::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(64) rows_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long nrows, int row_bytes, int NS, int do_store,
                                                  int lag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_mbar_init();
    const long long r0 = nrows * blockIdx.x / gridDim.x, r1 = nrows * (blockIdx.x + 1) / gridDim.x;
    const int my = (int)(r1 - r0);
    auto issue = [&](int it) {
      const int s = it % NS;
      ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)row_bytes);
      ptx::bulk_load_1d(buf + (size_t)s * row_bytes, in + (r0 + it) * row_bytes, (uint32_t)row_bytes, &full[s]);
    };
    for (int it = 0; it < NS && it < my; ++it) issue(it);
    for (int it = 0; it < my; ++it) {
      const int s = it % NS;
      ptx::mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
      if (do_store) {
        bulk_store_1d(out + (r0 + it) * row_bytes, buf + (size_t)s * row_bytes, (uint32_t)row_bytes);
        ptx::tma_store_commit();
        // the stage refilled now was stored `lag` iterations ago: wait until at most lag-1 younger stores are still reading
        if (lag <= 1) ptx::tma_store_wait_read0();
        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      const int nxt = it + NS - (do_store && lag > 1 ? 1 : 0);
      if (nxt >= NS && nxt < my && (do_store && lag > 1 ? true : true)) {
        if (!(do_store && lag > 1)) { if (it + NS < my) issue(it + NS); }
        else if (it >= 1 && it - 1 + NS < my) issue(it - 1 + NS);
      }
    }
    if (do_store) ptx::tma_store_wait_all0();
  }
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256, H = argc > 2 ? atoi(argv[2]) : 128, W = argc > 3 ? atoi(argv[3]) : 128, C = argc > 4 ? atoi(argv[4]) : 24;
  const size_t n = (size_t)B * H * W * C;
  float *in, *out;
  CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, n * 4));
  CK(cudaMemset(in, 1, n * 4)); CK(cudaMemset(out, 0, n * 4));
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  auto make = [&](float* base, int bh, int bw, int bc = 0) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)(bc ? bc : C), (cuuint32_t)bw, (cuuint32_t)bh, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  CUtensorMap tm_in = make(in, ITH, ITW), tm_out = make(out, TH, TW);
  const int CP = ((C / 4) | 1) * 4;
  CUtensorMap tm_in_pad = make(in, ITH, ITW, CP), tm_out_pad = make(out, TH, TW, CP);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, double bytes, auto&& launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    printf("%-34s %8.1f us  %7.1f GB/s\n", name, 1e3 * ms / iters, bytes / (ms / iters * 1e-3) / 1e9);
  };
  const double rd = (double)n * 4;
  timeit("ldg copy (read+write)", 2 * rd, [&] { ldg_copy<<<148 * 16, 256>>>((const float4*)in, (float4*)out, (long long)(n / 4)); });
  CK(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int stage_bytes = (ITH * ITW * C * 4 + 127) / 128 * 128;
  for (int ctas = 1; ctas <= 4; ctas *= 2)
    for (int NS : {2, 4, 6}) {
      const size_t sm = 1024 + (size_t)NS * stage_bytes;
      if (sm * ctas > 220 * 1024) continue;
      char name[96];
      snprintf(name, sizeof name, "tile loads   ctas/SM=%d NS=%d", ctas, NS);
      timeit(name, rd, [&] { tile_kernel<<<148 * ctas, 64, sm>>>(tm_in, tm_out, B, H, W, C, NS, 0, C); });
      snprintf(name, sizeof name, "tile ld+st   ctas/SM=%d NS=%d", ctas, NS);
      timeit(name, 2 * rd, [&] { tile_kernel<<<148 * ctas, 64, sm>>>(tm_in, tm_out, B, H, W, C, NS, 1, C); });
      const size_t smp = 1024 + (size_t)NS * ((ITH * ITW * CP * 4 + 127) / 128 * 128);
      if (smp * ctas <= 220 * 1024) {
        snprintf(name, sizeof name, "tile ld+st(pad-box st) c=%d NS=%d", ctas, NS);
        timeit(name, 2 * rd, [&] { tile_kernel<<<148 * ctas, 64, sm>>>(tm_in, tm_out_pad, B, H, W, C, NS, 1, C); });
        snprintf(name, sizeof name, "tile ld+st(pad-box ld+st) c=%d NS=%d", ctas, NS);
        timeit(name, 2 * rd, [&] { tile_kernel<<<148 * ctas, 64, smp>>>(tm_in_pad, tm_out_pad, B, H, W, C, NS, 1, CP); });
      }
    }
  const int row_bytes = W * C * 4;
  for (int ctas = 1; ctas <= 4; ctas *= 2)
    for (int NS : {2, 4, 8}) {
      const size_t sm = 1024 + (size_t)NS * row_bytes;
      if (sm * ctas > 220 * 1024) continue;
      char name[96];
      snprintf(name, sizeof name, "row loads    ctas/SM=%d NS=%d", ctas, NS);
      timeit(name, rd, [&] { rows_kernel<<<148 * ctas, 64, sm>>>((const uint8_t*)in, (uint8_t*)out, (long long)B * H, row_bytes, NS, 0, 1); });
      snprintf(name, sizeof name, "row ld+st    ctas/SM=%d NS=%d", ctas, NS);
      timeit(name, 2 * rd, [&] { rows_kernel<<<148 * ctas, 64, sm>>>((const uint8_t*)in, (uint8_t*)out, (long long)B * H, row_bytes, NS, 1, 1); });
      snprintf(name, sizeof name, "row ld+st(lag2) ctas/SM=%d NS=%d", ctas, NS);
      timeit(name, 2 * rd, [&] { rows_kernel<<<148 * ctas, 64, sm>>>((const uint8_t*)in, (uint8_t*)out, (long long)B * H, row_bytes, NS, 1, 2); });
    }
  return 0;
}
