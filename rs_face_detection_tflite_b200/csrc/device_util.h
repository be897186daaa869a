// device_util.h -- small RAII helpers for device / pinned buffers and frame staging.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "fdl_status.h"

namespace fdl {

// Every ABI call selects its handle's device; this guard puts the calling thread's current device back on exit, so a
// host process that drives several GPUs (its own CUDA work, torch, other handles) never sees its device changed by a call.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// CUDA events that are destroyed on every exit path.
struct EventSet {
  cudaEvent_t* ev = nullptr;
  int n = 0;
  ~EventSet() { for (int i = 0; i < n; ++i) if (ev[i]) cudaEventDestroy(ev[i]); delete[] ev; }
  cudaError_t create(int count) {
    ev = new cudaEvent_t[(size_t)count]();
    n = count;
    for (int i = 0; i < count; ++i) { cudaError_t e = cudaEventCreate(&ev[i]); if (e != cudaSuccess) return e; }
    return cudaSuccess;
  }
};

// Grow-only device buffer.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) { cudaDeviceSynchronize(); cudaFree(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
};

// Grow-only pinned host buffer.
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) { cudaDeviceSynchronize(); cudaFreeHost(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault);
    if (e == cudaSuccess) cap = n;
    return e;
  }
};

// Copies `n` equally-sized images (host or device) into one contiguous device buffer
// [n, height, width*3] on `stream`.  Returns FDL_OK or an error code (message set).
// When `direct` is given and the images are device-resident, tightly packed and contiguous, no copy is
// made: *direct receives the caller's pointer (which must stay valid until the work is collected).
inline int stage_frames(const fdl_image* images, int n, DevBuf<uint8_t>* dst, cudaStream_t stream, int* w_out, int* h_out,
                        const uint8_t** direct = nullptr, bool host_zero_copy = false, bool* used_host = nullptr) {
  if (!images || n <= 0) return set_error(FDL_ERR_INVALID, "no images given");
  const int w = images[0].width, h = images[0].height;
  if (w <= 0 || h <= 0) return set_error(FDL_ERR_INVALID, "image has non-positive size");
  const size_t row = (size_t)w * 3, frame = row * (size_t)h;
  for (int i = 0; i < n; ++i) {
    if (!images[i].data) return set_error(FDL_ERR_INVALID, "image data pointer is null");
    if (images[i].width != w || images[i].height != h) return set_error(FDL_ERR_INVALID, "all images of a batch must have the same size");
    if (images[i].row_stride < 0) return set_error(FDL_ERR_INVALID, "negative row_stride (bottom-up images are not supported: pass a top-down copy)");
    if (images[i].row_stride != 0 && (size_t)images[i].row_stride < row) return set_error(FDL_ERR_INVALID, "row_stride smaller than width*3");
  }
  if (direct) {
    bool contiguous = true;
    for (int i = 0; i < n && contiguous; ++i)
      contiguous = images[i].mem == FDL_MEM_DEVICE && (images[i].row_stride == 0 || (size_t)images[i].row_stride == row) &&
                   images[i].data == images[0].data + (size_t)i * frame;
    if (contiguous) { *direct = images[0].data; *w_out = w; *h_out = h; return FDL_OK; }
    if (host_zero_copy) {
      // pinned (page-locked, mapped) host frames: let the kernels read them in place over PCIe instead of
      // copying whole frames -- the letterbox touches 27 % of a 1080p frame, the ROI warps a few hundred KB
      bool hc = true;
      for (int i = 0; i < n && hc; ++i)
        hc = images[i].mem == FDL_MEM_HOST && (images[i].row_stride == 0 || (size_t)images[i].row_stride == row) &&
             images[i].data == images[0].data + (size_t)i * frame;
      if (hc) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, images[0].data) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
          *direct = static_cast<const uint8_t*>(at.devicePointer);
          if (used_host) *used_host = true;
          *w_out = w; *h_out = h;
          return FDL_OK;
        }
        cudaGetLastError();
      }
    }
  }
  FDL_CUDA_TRY(dst->reserve(frame * (size_t)n));
  // coalesce runs of frames that are contiguous in the caller's memory into one copy
  int i = 0;
  while (i < n) {
    const size_t stride = images[i].row_stride ? (size_t)images[i].row_stride : row;
    const cudaMemcpyKind kind = images[i].mem == FDL_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int j = i + 1;
    if (stride == row) {
      while (j < n && images[j].mem == images[i].mem && (images[j].row_stride == 0 || (size_t)images[j].row_stride == row) &&
             images[j].data == images[i].data + (size_t)(j - i) * frame)
        ++j;
      FDL_CUDA_TRY(cudaMemcpyAsync(dst->p + (size_t)i * frame, images[i].data, frame * (size_t)(j - i), kind, stream));
    } else {
      FDL_CUDA_TRY(cudaMemcpy2DAsync(dst->p + (size_t)i * frame, row, images[i].data, stride, row, (size_t)h, kind, stream));
    }
    i = j;
  }
  if (direct) *direct = dst->p;
  *w_out = w; *h_out = h;
  return FDL_OK;
}

}  // namespace fdl
