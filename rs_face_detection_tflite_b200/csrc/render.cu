// render.cu -- the drawing half of SURVEY.md 8f rank 4 on the device: render.rs:361-479 `render_to_image` (annotations painted over
// the frame, later ones over earlier ones, no blending) with the pixel rules of the `imageproc` crate the reference calls
// (draw_filled_rect_mut / draw_hollow_rect_mut clip to the image; draw_line_segment_mut walks BresenhamLineIter in f32).  Painted
// pixel sets are held to the reference's own rendered assets (tests/test_gpu_render.py against tests/golden/).
//
// Two launches: every primitive's thread marks the pixels it covers with its index (atomicMax: the LAST primitive that touches a
// pixel owns it -- the sequential overwrite order of the reference, without a race), then one thread per pixel composes RGBA.
#include <cuda_runtime.h>

#include <climits>
#include <cstring>
#include <string>

#include "device_util.h"
#include "fdl_status.h"

namespace fdl {
void count_launch();

namespace {

// f64 -> u32 / i32 as Rust's `as` casts: truncate toward zero, saturate, NaN -> 0
__device__ __forceinline__ unsigned as_u32(double v) { return v != v ? 0u : (v <= 0.0 ? 0u : (v >= 4294967295.0 ? 4294967295u : (unsigned)v)); }
__device__ __forceinline__ int as_i32(double v) { return v != v ? 0 : (v <= -2147483648.0 ? INT_MIN : (v >= 2147483647.0 ? INT_MAX : (int)v)); }

__device__ __forceinline__ void mark(int* owner, int W, int H, long long x, long long y, int tag) {
  if (x >= 0 && x < W && y >= 0 && y < H) atomicMax(&owner[y * W + x], tag);
}

// imageproc draw_line_segment_mut: BresenhamLineIter on (f32, f32) end points (here whole numbers)
__device__ void draw_line(int* owner, int W, int H, float x0, float y0, float x1, float y1, int tag) {
  const bool steep = fabsf(y1 - y0) > fabsf(x1 - x0);
  if (steep) { float t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
  if (x0 > x1) { float t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
  const float dx = x1 - x0, dy = fabsf(y1 - y0);
  float err = dx / 2.0f;
  long long x = (long long)x0, y = (long long)y0;
  const long long end_x = (long long)x1, y_step = y0 < y1 ? 1 : -1;
  for (; x <= end_x; ++x) {                       // (the host bounds the coordinates, so the walk is short)
    if (steep) mark(owner, W, H, y, x, tag); else mark(owner, W, H, x, y, tag);
    err -= dy;
    if (err < 0.0f) { y += y_step; err += dx; }
  }
}

__device__ void fill_rect(int* owner, int W, int H, long long left, long long top, long long rw, long long rh, int tag) {
  const long long x0 = left < 0 ? 0 : left, y0 = top < 0 ? 0 : top;
  const long long x1 = left + rw > W ? W : left + rw, y1 = top + rh > H ? H : top + rh;
  for (long long y = y0; y < y1; ++y)
    for (long long x = x0; x < x1; ++x) atomicMax(&owner[y * W + x], tag);
}

__global__ void render_mark_kernel(const fdl_primitive* __restrict__ prims, int n, int W, int H, int* __restrict__ owner, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const fdl_primitive p = prims[i];
  const double sx = p.normalized ? (double)W : 1.0, sy = p.normalized ? (double)H : 1.0;
  const int tag = i + 1;
  const unsigned thick = as_u32(p.thickness);
  if (p.kind == FDL_PRIM_POINT) {
    const unsigned w = thick / 2 > 1 ? thick / 2 : 1;
    const unsigned x = as_u32(p.a * sx), y = as_u32(p.b * sy);
    fill_rect(owner, W, H, (long long)(int)(x - w), (long long)(int)(y - w), 2ll * w, 2ll * w, tag);     // u32 arithmetic, then `as i32`
  } else if (p.kind == FDL_PRIM_LINE) {
    draw_line(owner, W, H, (float)as_i32(p.a * sx), (float)as_i32(p.b * sy), (float)as_i32(p.c * sx), (float)as_i32(p.d * sy), tag);
  } else if (p.kind == FDL_PRIM_RECT || p.kind == FDL_PRIM_FILLED_RECT) {
    const double l = p.a * sx, t = p.b * sy, r = p.c * sx, b = p.d * sy;
    const long long left = as_i32(l), top = as_i32(t), rw = as_u32(r - l), rh = as_u32(b - t);
    if (rw == 0 || rh == 0) { atomicExch(bad, i + 1); return; }     // imageproc's Rect::of_size panics on an empty size
    if (p.kind == FDL_PRIM_FILLED_RECT) { fill_rect(owner, W, H, left, top, rw, rh, tag); return; }
    const float fl = (float)left, ft = (float)top, fr = (float)(left + rw - 1), fb = (float)(top + rh - 1);
    draw_line(owner, W, H, fl, ft, fr, ft, tag);
    draw_line(owner, W, H, fl, fb, fr, fb, tag);
    draw_line(owner, W, H, fl, ft, fl, fb, tag);
    draw_line(owner, W, H, fr, ft, fr, fb, tag);
  } else {
    atomicExch(bad, i + 1);
  }
}

__global__ void render_compose_kernel(const uint8_t* __restrict__ rgb, long long row_stride, const fdl_primitive* __restrict__ prims,
                                      const int* __restrict__ owner, int W, int H, uint8_t* __restrict__ out_rgba) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)W * H) return;
  const int y = (int)(i / W), x = (int)(i - (long long)y * W);
  const int o = owner[i];
  uchar4 px;
  if (o > 0) {
    const fdl_primitive& p = prims[o - 1];
    px = make_uchar4(p.r, p.g, p.b_, p.alpha);
  } else {
    const uint8_t* s = rgb + (long long)y * row_stride + 3ll * x;
    px = make_uchar4(s[0], s[1], s[2], 255);          // DynamicImage::to_rgba8 of an RGB image
  }
  reinterpret_cast<uchar4*>(out_rgba)[i] = px;
}

}  // namespace
}  // namespace fdl

using namespace fdl;

extern "C" {

int fdl_render_to_image(int device, const fdl_image* image, const fdl_primitive* primitives, int n, uint8_t* out_rgba, size_t cap, int out_mem) try {
  DeviceGuard _device_guard;
  if (!image || (n > 0 && !primitives) || n < 0 || !out_rgba) return set_error(FDL_ERR_INVALID, "bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(FDL_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)"); }
  if (device < 0 || device >= ndev) return set_error(FDL_ERR_INVALID, "device index out of range");
  FDL_CUDA_TRY(cudaSetDevice(device));
  const int W = image->width, H = image->height;
  if (W <= 0 || H <= 0) return set_error(FDL_ERR_INVALID, "image has non-positive size");
  // the walk of a line visits every column between its end points, inside the image or not: bound the coordinates (the reference
  // would simply take that long)
  for (int i = 0; i < n; ++i) {
    const fdl_primitive& q = primitives[i];
    const double sx = q.normalized ? (double)W : 1.0, sy = q.normalized ? (double)H : 1.0;
    const double v[4] = {q.a * sx, q.b * sy, q.c * sx, q.d * sy};
    for (double c : v)
      if (!(c > -1.0e6 && c < 1.0e6)) return set_error(FDL_ERR_INVALID, "primitive " + std::to_string(i) + ": coordinate out of range");
  }
  const size_t bytes = (size_t)W * H * 4;
  if (cap < bytes) return set_error(FDL_ERR_CAPACITY, "output buffer too small: width * height * 4 bytes (RGBA) needed");
  DevBuf<uint8_t> frame, out;
  DevBuf<int> owner, bad;
  DevBuf<fdl_primitive> prims;
  int w, h;
  const uint8_t* fptr = nullptr;
  int rc = stage_frames(image, 1, &frame, 0, &w, &h, &fptr);
  if (rc) return rc;
  FDL_CUDA_TRY(owner.reserve((size_t)W * H));
  FDL_CUDA_TRY(bad.reserve(1));
  FDL_CUDA_TRY(prims.reserve((size_t)(n > 0 ? n : 1)));
  FDL_CUDA_TRY(cudaMemsetAsync(owner.p, 0, (size_t)W * H * sizeof(int), 0));
  FDL_CUDA_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), 0));
  if (n > 0) {
    FDL_CUDA_TRY(cudaMemcpyAsync(prims.p, primitives, (size_t)n * sizeof(fdl_primitive), cudaMemcpyHostToDevice, 0));
    render_mark_kernel<<<(n + 127) / 128, 128>>>(prims.p, n, W, H, owner.p, bad.p);
    count_launch();
    FDL_CUDA_TRY(cudaGetLastError());
  }
  uint8_t* dst = out_rgba;
  if (out_mem != FDL_MEM_DEVICE) { FDL_CUDA_TRY(out.reserve(bytes)); dst = out.p; }
  render_compose_kernel<<<(unsigned)(((long long)W * H + 255) / 256), 256>>>(fptr, (long long)W * 3, prims.p, owner.p, W, H, dst);
  count_launch();
  FDL_CUDA_TRY(cudaGetLastError());
  int hbad = 0;
  FDL_CUDA_TRY(cudaMemcpy(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (out_mem != FDL_MEM_DEVICE) FDL_CUDA_TRY(cudaMemcpy(out_rgba, out.p, bytes, cudaMemcpyDeviceToHost));
  if (hbad) return set_error(FDL_ERR_INVALID, "primitive " + std::to_string(hbad - 1) + ": unknown kind, or a rectangle of zero width or height");
  return FDL_OK;
} FDL_ABI_CATCH

}  // extern "C"
