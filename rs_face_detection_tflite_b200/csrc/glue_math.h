// glue_math.h -- scalar math of the reference's Rust glue and of the OpenCV ops it calls, written
// once as host+device inline functions.  The CUDA kernels in prepost_kernels.cu call these on the
// device; tests/hostcheck/ compiles the very same header with g++ so the arithmetic can be checked
// against the oracle on a CPU-only box (test infrastructure only -- the product library never runs
// them on the host).
//
// Every function cites the reference lines it follows (relative to
// /root/reference/src/face_detection_lite/).  f32/f64 casts follow the Rust types.  This header is
// compiled with FMA contraction disabled (-fmad=false / -ffp-contract=off): the reference's scalar
// Rust code never fuses a multiply with an add.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/fdl.h"

#if defined(__CUDACC__)
#define FDL_HD __host__ __device__ __forceinline__
#else
#define FDL_HD inline
#endif

namespace fdl {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// Rust `x as i32` for f64: truncation toward zero, saturating, NaN -> 0.
FDL_HD int f64_as_i32(double v) {
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return (int)v;
}
FDL_HD int imin(int a, int b) { return a < b ? a : b; }
FDL_HD int imax(int a, int b) { return a > b ? a : b; }
FDL_HD double dmin(double a, double b) { return a < b ? a : b; }
FDL_HD double dmax(double a, double b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------
// types.rs
// ------------------------------------------------------------------------------------------------
// Rect::scaled (types.rs:62-77)
FDL_HD fdl_rect rect_scaled(const fdl_rect& r, double sw, double sh, bool normalize) {
  if ((r.normalized != 0) == normalize) return r;
  double sx = normalize ? 1.0 / sw : sw, sy = normalize ? 1.0 / sh : sh;
  fdl_rect o = r;
  o.x_center = r.x_center * sx; o.y_center = r.y_center * sy;
  o.width = r.width * sx; o.height = r.height * sy;
  o.normalized = normalize ? 1 : 0;
  return o;
}
// Rect::points (types.rs:80-96): TL, TR, BR, BL, rotated about the centre.
FDL_HD void rect_points(const fdl_rect& r, double px[4], double py[4]) {
  double x = r.x_center, y = r.y_center, w = r.width / 2.0, h = r.height / 2.0;
  px[0] = x - w; py[0] = y - h;
  px[1] = x + w; py[1] = y - h;
  px[2] = x + w; py[2] = y + h;
  px[3] = x - w; py[3] = y + h;
  if (r.rotation != 0.0) {
    double s = sin(r.rotation), c = cos(r.rotation);
    for (int i = 0; i < 4; ++i) {
      double dx = px[i] - x, dy = py[i] - y;
      px[i] = x + dx * c - dy * s;
      py[i] = y + dx * s + dy * c;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// OpenCV getPerspectiveTransform + invert (transform.rs:222)
// ------------------------------------------------------------------------------------------------
// Solves the 8x8 system OpenCV builds (src (x,y) -> dst (u,v)) by f64 Gaussian elimination with
// partial pivoting.  The reference asks for DECOMP_SVD (it passes INTER_LINEAR == 1 as the solve
// method); both agree far below the 1/32-px quantisation of the warp.  Returns false if singular.
FDL_HD bool perspective_transform(const float sx[4], const float sy[4], const float dx[4], const float dy[4], double M[9]) {
  double a[8][9];
  for (int i = 0; i < 4; ++i) {
    double x = sx[i], y = sy[i], u = dx[i], v = dy[i];
    double* r0 = a[i];
    double* r1 = a[i + 4];
    r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -x * u; r0[7] = -y * u; r0[8] = u;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -x * v; r1[7] = -y * v; r1[8] = v;
  }
  for (int k = 0; k < 8; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < 8; ++i) { double t = fabs(a[i][k]); if (t > best) { best = t; p = i; } }
    if (!(best > 0.0)) return false;
    if (p != k) for (int j = 0; j < 9; ++j) { double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
    for (int i = k + 1; i < 8; ++i) {
      double f = a[i][k] / a[k][k];
      for (int j = k; j < 9; ++j) a[i][j] -= f * a[k][j];
    }
  }
  double xs[8];
  for (int k = 7; k >= 0; --k) {
    double s = a[k][8];
    for (int j = k + 1; j < 8; ++j) s -= a[k][j] * xs[j];
    xs[k] = s / a[k][k];
  }
  for (int i = 0; i < 8; ++i) M[i] = xs[i];
  M[8] = 1.0;
  return true;
}
// cv::invert of a 3x3 f64 matrix (closed-form cofactors).
FDL_HD bool invert3x3(const double m[9], double t[9]) {
  double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (d == 0.0) return false;
  d = 1.0 / d;
  t[0] = (m[4] * m[8] - m[5] * m[7]) * d;
  t[1] = (m[2] * m[7] - m[1] * m[8]) * d;
  t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * d;
  t[4] = (m[0] * m[8] - m[2] * m[6]) * d;
  t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * d;
  t[7] = (m[1] * m[6] - m[0] * m[7]) * d;
  t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
  return true;
}

// ------------------------------------------------------------------------------------------------
// image_to_tensor (transform.rs:188-309): everything that does not depend on the pixel
// ------------------------------------------------------------------------------------------------
struct I2TParams {
  double Mi[9];          // inverse perspective matrix (dst -> src), cv::warpPerspective's working matrix
  double pad[4];         // ImageTensor.padding (left, top, right, bottom)
  int32_t src_w, src_h;  // frame size
  int32_t warp_w, warp_h;   // size of the warpPerspective output
  int32_t pad_h, pad_v;     // copyMakeBorder amounts (0 when that stage is skipped)
  int32_t r1_w, r1_h;       // size after the first resize (== bordered size when the stage is skipped / identity)
  int32_t has_r1;           // copyMakeBorder + resize(new_width,new_height) stage present (transform.rs:251-274)
  int32_t has_r2;           // final resize to the output size present (keep_aspect_ratio, :276-280)
  int32_t out_w, out_h;
  int32_t flip;
  int32_t valid;            // 0: the reference would have returned Err / OpenCV would have thrown
  int32_t frame;            // index of the source frame in the batch
  int32_t _pad;
  float scale, offset;      // tensor = f32(f64(px) * scale_d + offset_d); kept in f64 below
  double range_min, range_max;
};

FDL_HD void i2t_setup(const fdl_rect* roi_or_null, int img_w, int img_h, int out_w, int out_h, bool keep_aspect,
                      double range_min, double range_max, bool flip, int frame, I2TParams* P) {
  fdl_rect roi;
  if (roi_or_null) roi = *roi_or_null;
  else { roi.x_center = 0.5; roi.y_center = 0.5; roi.width = 1.0; roi.height = 1.0; roi.rotation = 0.0; roi.normalized = 1; }
  roi = rect_scaled(roi, (double)img_w, (double)img_h, false);                       // :199
  int width, height;
  if (keep_aspect) { width = f64_as_i32((double)f64_as_i32(roi.width)); height = f64_as_i32((double)f64_as_i32(roi.height)); }  // :203-207, types.rs:52-59
  else { width = out_w; height = out_h; }
  P->src_w = img_w; P->src_h = img_h; P->warp_w = width; P->warp_h = height;
  P->out_w = out_w; P->out_h = out_h; P->flip = flip ? 1 : 0; P->frame = frame;
  P->range_min = range_min; P->range_max = range_max;
  P->pad_h = P->pad_v = 0; P->has_r1 = 0; P->has_r2 = keep_aspect ? 1 : 0; P->r1_w = width; P->r1_h = height;
  P->pad[0] = P->pad[1] = P->pad[2] = P->pad[3] = 0.0;
  P->valid = 1;
  if (width <= 0 || height <= 0 || out_w <= 0 || out_h <= 0 || img_w <= 0 || img_h <= 0) { P->valid = 0; return; }
  double px[4], py[4];
  rect_points(roi, px, py);
  float sx[4], sy[4];
  for (int i = 0; i < 4; ++i) { sx[i] = (float)px[i]; sy[i] = (float)py[i]; }          // :210-213
  float dx[4] = {0.f, (float)width, (float)width, 0.f}, dy[4] = {0.f, 0.f, (float)height, (float)height};  // :216-220
  double M[9];
  if (!perspective_transform(sx, sy, dx, dy, M) || !invert3x3(M, P->Mi)) { P->valid = 0; return; }
  if (keep_aspect) {
    double out_aspect = (double)(out_h / out_w);                                      // integer division (sic) :240
    double roi_aspect = roi.height / roi.width;
    int new_w = f64_as_i32(roi.width), new_h = f64_as_i32(roi.height);
    double pad_x = 0.0, pad_y = 0.0;
    if (out_aspect > roi_aspect) { new_h = f64_as_i32(roi.width * out_aspect); pad_y = (1.0 - roi_aspect / out_aspect) / 2.0; }
    else { new_w = f64_as_i32(roi.height / out_aspect); pad_x = (1.0 - out_aspect / roi_aspect) / 2.0; }
    if (new_w != f64_as_i32(roi.width) || new_h != f64_as_i32(roi.height)) {
      P->pad_h = f64_as_i32(pad_x * (double)new_w);
      P->pad_v = f64_as_i32(pad_y * (double)new_h);
      P->has_r1 = 1; P->r1_w = new_w; P->r1_h = new_h;
      if (new_w <= 0 || new_h <= 0 || P->pad_h < 0 || P->pad_v < 0) { P->valid = 0; return; }
    }
    P->pad[0] = pad_x; P->pad[1] = pad_y; P->pad[2] = pad_x; P->pad[3] = pad_y;
  }
}

// ------------------------------------------------------------------------------------------------
// per-pixel evaluation of the OpenCV chain warpPerspective -> copyMakeBorder -> resize -> resize
// ------------------------------------------------------------------------------------------------
struct Px3 { int r, g, b; };

// Where source pixels come from: the frame in (global / mapped host) memory, optionally shadowed by a staged copy
// of the rectangle [tx0,tx1] x [ty0,ty1] (a kernel's shared-memory tile; row r of the tile holds the frame bytes
// [tsb, tsb + tpitch) of frame row ty0 + r).  Pixels outside the staged rectangle fall through to `img`.
struct ImgSrc {
  const uint8_t* img;
  long long stride;          // bytes between frame rows
  const uint8_t* tile;       // nullptr: no staged copy
  int tx0, ty0, tx1, ty1;    // staged pixel rectangle (inclusive)
  int tpitch, tsb;
};
FDL_HD ImgSrc img_src(const uint8_t* img, long long stride) {
  ImgSrc s; s.img = img; s.stride = stride; s.tile = nullptr; s.tx0 = s.ty0 = 0; s.tx1 = s.ty1 = -1; s.tpitch = 0; s.tsb = 0;
  return s;
}
// One u8 RGB pixel of the source frame.
FDL_HD Px3 load_px(const ImgSrc& s, int x, int y) {
  const uint8_t* p;
  if (s.tile && x >= s.tx0 && x <= s.tx1 && y >= s.ty0 && y <= s.ty1) p = s.tile + (long long)(y - s.ty0) * s.tpitch + (3 * x - s.tsb);
  else p = s.img + (long long)y * s.stride + 3 * x;
  Px3 o; o.r = p[0]; o.g = p[1]; o.b = p[2];
  return o;
}

// cv::warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0) at destination pixel (x,y): SURVEY.md B.2.
// source position of destination pixel (x,y) in 1/32-pixel fixed point: top-left tap (sx, sy) and the fractions (ax, ay) / 32
FDL_HD void warp_coords(const I2TParams& P, int x, int y, int* sx_out, int* sy_out, int* ax_out, int* ay_out) {
  const double* Mi = P.Mi;
  double W = Mi[6] * x + Mi[7] * y + Mi[8];
  W = W != 0.0 ? 32.0 / W : 0.0;
  double fX = (Mi[0] * x + Mi[1] * y + Mi[2]) * W;
  double fY = (Mi[3] * x + Mi[4] * y + Mi[5]) * W;
  fX = dmax(-2147483648.0, dmin(2147483647.0, fX));
  fY = dmax(-2147483648.0, dmin(2147483647.0, fY));
  int X = (int)rint(fX), Y = (int)rint(fY);            // clamped to the int32 range above: the conversion is exact
  int sxl = X >> 5, syl = Y >> 5;
  *sx_out = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);   // saturate_cast<short>
  *sy_out = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
  *ax_out = X & 31; *ay_out = Y & 31;
}

FDL_HD Px3 warp_px(const I2TParams& P, const ImgSrc& src, int x, int y) {
  int sx, sy, ax, ay;
  warp_coords(P, x, y, &sx, &sy, &ax, &ay);
  // OpenCV's bilinear table: float32 products of (1-fy),(fy) x (1-fx),(fx), scaled by 32768 and rounded
  float fx = (float)ax * (1.f / 32.f), fy = (float)ay * (1.f / 32.f);
  int w00 = (int)rintf((1.f - fy) * (1.f - fx) * 32768.f);
  int w01 = (int)rintf((1.f - fy) * fx * 32768.f);
  int w10 = (int)rintf(fy * (1.f - fx) * 32768.f);
  int w11 = (int)rintf(fy * fx * 32768.f);
  int r = 0, g = 0, b = 0;
  bool x0 = sx >= 0 && sx < P.src_w, x1 = sx + 1 >= 0 && sx + 1 < P.src_w;
  bool y0 = sy >= 0 && sy < P.src_h, y1 = sy + 1 >= 0 && sy + 1 < P.src_h;
  if (w00 && x0 && y0) { Px3 p = load_px(src, sx, sy); r += p.r * w00; g += p.g * w00; b += p.b * w00; }
  if (w01 && x1 && y0) { Px3 p = load_px(src, sx + 1, sy); r += p.r * w01; g += p.g * w01; b += p.b * w01; }
  if (w10 && x0 && y1) { Px3 p = load_px(src, sx, sy + 1); r += p.r * w10; g += p.g * w10; b += p.b * w10; }
  if (w11 && x1 && y1) { Px3 p = load_px(src, sx + 1, sy + 1); r += p.r * w11; g += p.g * w11; b += p.b * w11; }
  Px3 o;
  o.r = (r + (1 << 14)) >> 15; o.g = (g + (1 << 14)) >> 15; o.b = (b + (1 << 14)) >> 15;
  return o;
}

// ------------------------------------------------------------------------------------------------
// Zero-copy host frames: which frame bytes a warp can touch (roi_fill_kernel / eye_split_kernel, prepost_kernels.cu; checked on
// the host by tests/hostcheck against every tap of warp_px).
// Source region of a warp in frame pixels: the bounding rectangle of its taps and, when `quad` is set, the convex quadrilateral
// (corners of warp space through the inverse matrix, in polygon order) the samples themselves lie in.
struct SrcBox { int x0, y0, x1, y1; int quad, _pad; double qx[4], qy[4]; };

FDL_HD SrcBox warp_src_box(const I2TParams& P) {
  // the corners of warp space through the inverse matrix: every tap of the warp lies in [floor(lo) - 1, ceil(hi) + 2]
  SrcBox b; b.x0 = 0; b.y0 = 0; b.x1 = -1; b.y1 = -1; b.quad = 0; b._pad = 0;
  for (int c = 0; c < 4; ++c) b.qx[c] = b.qy[c] = 0.0;
  if (P.valid != 1) return b;
  double lox = 1e30, loy = 1e30, hix = -1e30, hiy = -1e30;
  for (int c = 0; c < 4; ++c) {
    const int xr = (c == 1 || c == 2), yr = (c >= 2);              // polygon order: (0,0) (w-1,0) (w-1,h-1) (0,h-1)
    const double x = xr ? (double)(P.warp_w - 1) : 0.0, y = yr ? (double)(P.warp_h - 1) : 0.0;
    const double w = P.Mi[6] * x + P.Mi[7] * y + P.Mi[8];
    if (!(w > 1e-6)) return b;
    const double sx = (P.Mi[0] * x + P.Mi[1] * y + P.Mi[2]) / w, sy = (P.Mi[3] * x + P.Mi[4] * y + P.Mi[5]) / w;
    b.qx[c] = sx; b.qy[c] = sy;
    lox = dmin(lox, sx); hix = dmax(hix, sx); loy = dmin(loy, sy); hiy = dmax(hiy, sy);
  }
  if (!(hix - lox < 1e5 && hiy - loy < 1e5 && lox > -1e6 && loy > -1e6)) return b;
  b.x0 = (int)floor(lox) - 1; b.y0 = (int)floor(loy) - 1; b.x1 = (int)ceil(hix) + 2; b.y1 = (int)ceil(hiy) + 2;
  // every sample of the warp stage (whatever border / resize stages follow it) lies inside the image of warp space, a convex
  // quadrilateral
  b.quad = 1;
  return b;
}

// x extent of the quadrilateral over the rows [ya, yb]: false when it does not reach the band
FDL_HD bool quad_span(const SrcBox& b, double ya, double yb, double* xlo, double* xhi) {
  double lo = 1e30, hi = -1e30;
  for (int e = 0; e < 4; ++e) {
    const double x0 = b.qx[e], y0 = b.qy[e], x1 = b.qx[(e + 1) & 3], y1 = b.qy[(e + 1) & 3];
    if (y0 >= ya && y0 <= yb) { lo = dmin(lo, x0); hi = dmax(hi, x0); }
    if ((y0 - ya) * (y1 - ya) < 0.0) { const double x = x0 + (ya - y0) / (y1 - y0) * (x1 - x0); lo = dmin(lo, x); hi = dmax(hi, x); }
    if ((y0 - yb) * (y1 - yb) < 0.0) { const double x = x0 + (yb - y0) / (y1 - y0) * (x1 - x0); lo = dmin(lo, x); hi = dmax(hi, x); }
  }
  *xlo = lo; *xhi = hi;
  return hi >= lo;
}

// signed distance-like test: is (x, y) inside the convex quadrilateral, at least `d` pixels from every edge?
FDL_HD bool quad_contains(const SrcBox& b, double x, double y, double d) {
  int pos = 0, neg = 0;
  for (int e = 0; e < 4; ++e) {
    const double x0 = b.qx[e], y0 = b.qy[e], ex = b.qx[(e + 1) & 3] - x0, ey = b.qy[(e + 1) & 3] - y0;
    const double len = sqrt(ex * ex + ey * ey);
    if (!(len > 1e-9)) return false;
    const double dist = (ex * (y - y0) - ey * (x - x0)) / len;     // > 0 on one side of the edge for every edge of a convex polygon
    if (dist >= d) ++pos; else if (dist <= -d) ++neg; else return false;
  }
  return pos == 4 || neg == 4;
}

// The region roi_fill_kernel stages for a face slot: the tap rectangle grown by margin m = pct % of its larger side + 2 px and
// clipped to the frame; `trim`: rows are copied only over the quadrilateral's span (roi_row_span).
FDL_HD SrcBox roi_stage_box(const I2TParams& P, int margin_pct, bool trim, int* margin) {
  SrcBox b = warp_src_box(P);
  *margin = 0;
  if (b.x1 >= b.x0) {
    const int m = (imax(b.x1 - b.x0, b.y1 - b.y0) * margin_pct) / 100 + 2;
    b.x0 = imax(b.x0 - m, 0); b.y0 = imax(b.y0 - m, 0); b.x1 = imin(b.x1 + m, P.src_w - 1); b.y1 = imin(b.y1 + m, P.src_h - 1);
    *margin = m;
    if (!trim) b.quad = 0;
  }
  return b;
}

// Columns [x0, x1] of frame row r that are staged: taps of samples with |sy - r| <= 1 (+ the fixed-point rounding of the warp
// coordinates and the margin), widened by the same -1 / +2 tap slack as the rectangle.  false: nothing of this row.
FDL_HD bool roi_row_span(const SrcBox& b, int r, int margin, int* x0, int* x1) {
  if (r < b.y0 || r > b.y1 || b.x1 < b.x0) return false;
  *x0 = b.x0; *x1 = b.x1;
  if (b.quad) {
    double lo, hi;
    if (!quad_span(b, (double)r - 1.25 - margin, (double)r + 1.25 + margin, &lo, &hi)) return false;
    *x0 = imax((int)floor(lo) - 1 - margin, b.x0); *x1 = imin((int)ceil(hi) + 2 + margin, b.x1);
  }
  return *x1 >= *x0;
}

// May a warp with source region `e` (warp_src_box of an eye slot) run on the staged copy of face region `f`?  Its tap rectangle
// must be inside the face's; when the face rows were trimmed, every sample of the eye warp must lie inside the face quadrilateral
// (then its taps are taps the face warp's own samples could have, which is what the per-row spans cover).
FDL_HD bool roi_stage_covers(const SrcBox& f, const SrcBox& e, int src_w, int src_h) {
  if (!(e.x1 >= e.x0 && f.x1 >= f.x0)) return false;
  if (!(imax(e.x0, 0) >= f.x0 && imax(e.y0, 0) >= f.y0 && imin(e.x1, src_w - 1) <= f.x1 && imin(e.y1, src_h - 1) <= f.y1)) return false;
  if (f.quad) {
    if (!e.quad) return false;
    for (int c = 0; c < 4; ++c) if (!quad_contains(f, e.qx[c], e.qy[c], 0.25)) return false;
  }
  return true;
}

// cv::resize(INTER_LINEAR) coefficients for one axis. clamp_frac: x axis.  `scale` = (double)sn / (double)dn
// (callers that evaluate many coordinates of one axis divide once).
FDL_HD void resize_coeff_scaled(int d, double scale, int sn, bool clamp_frac, int* s0, int* s1, int* c0, int* c1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  float fr = f - (float)s;
  if (clamp_frac) {
    if (s < 0) { fr = 0.f; s = 0; }
    if (s >= sn - 1) { fr = 0.f; s = sn - 1; }
    *s0 = s; *s1 = imin(s + 1, sn - 1);
  } else {
    *s0 = imin(imax(s, 0), sn - 1); *s1 = imin(imax(s + 1, 0), sn - 1);
  }
  *c0 = (int)rintf((1.f - fr) * 2048.f);
  *c1 = (int)rintf(fr * 2048.f);
}
FDL_HD void resize_coeff(int d, int dn, int sn, bool clamp_frac, int* s0, int* s1, int* c0, int* c1) {
  resize_coeff_scaled(d, (double)sn / (double)dn, sn, clamp_frac, s0, s1, c0, c1);
}

FDL_HD int resize_mix(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
  int h0 = p00 * a0 + p01 * a1, h1 = p10 * a0 + p11 * a1;
  int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// pixel of the copyMakeBorder'ed warp (constant 0 border)
FDL_HD Px3 border_px(const I2TParams& P, const ImgSrc& src, int x, int y) {
  int wx = x - P.pad_h, wy = y - P.pad_v;
  if (wx < 0 || wy < 0 || wx >= P.warp_w || wy >= P.warp_h) { Px3 z; z.r = z.g = z.b = 0; return z; }
  return warp_px(P, src, wx, wy);
}

// pixel of the image after the optional first resize stage (size r1_w x r1_h)
FDL_HD Px3 stage1_px(const I2TParams& P, const ImgSrc& src, int x, int y) {
  if (!P.has_r1) return warp_px(P, src, x, y);
  int bw = P.warp_w + 2 * P.pad_h, bh = P.warp_h + 2 * P.pad_v;
  if (bw == P.r1_w && bh == P.r1_h) return border_px(P, src, x, y);   // cv::resize to the same size is a copy
  int x0, x1, a0, a1, y0, y1, b0, b1;
  resize_coeff(x, P.r1_w, bw, true, &x0, &x1, &a0, &a1);
  resize_coeff(y, P.r1_h, bh, false, &y0, &y1, &b0, &b1);
  Px3 p00 = border_px(P, src, x0, y0), p01 = border_px(P, src, x1, y0);
  Px3 p10 = border_px(P, src, x0, y1), p11 = border_px(P, src, x1, y1);
  Px3 o;
  o.r = resize_mix(p00.r, p01.r, p10.r, p11.r, a0, a1, b0, b1);
  o.g = resize_mix(p00.g, p01.g, p10.g, p11.g, a0, a1, b0, b1);
  o.b = resize_mix(p00.b, p01.b, p10.b, p11.b, a0, a1, b0, b1);
  return o;
}

// final uint8 pixel of image_to_tensor's `roi_image` at output position (ox, oy) (after the flip)
FDL_HD Px3 i2t_pixel(const I2TParams& P, const ImgSrc& src, int ox, int oy) {
  int x = P.flip ? P.out_w - 1 - ox : ox;
  if (!P.has_r2) return warp_px(P, src, x, oy);
  if (P.r1_w == P.out_w && P.r1_h == P.out_h) return stage1_px(P, src, x, oy);
  int x0, x1, a0, a1, y0, y1, b0, b1;
  resize_coeff(x, P.out_w, P.r1_w, true, &x0, &x1, &a0, &a1);
  resize_coeff(oy, P.out_h, P.r1_h, false, &y0, &y1, &b0, &b1);
  Px3 p00 = stage1_px(P, src, x0, y0), p01 = stage1_px(P, src, x1, y0);
  Px3 p10 = stage1_px(P, src, x0, y1), p11 = stage1_px(P, src, x1, y1);
  Px3 o;
  o.r = resize_mix(p00.r, p01.r, p10.r, p11.r, a0, a1, b0, b1);
  o.g = resize_mix(p00.g, p01.g, p10.g, p11.g, a0, a1, b0, b1);
  o.b = resize_mix(p00.b, p01.b, p10.b, p11.b, a0, a1, b0, b1);
  return o;
}

// normalisation (transform.rs:298): f32( f64(px) * (max - min) / 255.0 + min )
FDL_HD float i2t_normalise(int px, double range_min, double range_max) {
  return (float)((double)px * (range_max - range_min) / 255.0 + range_min);
}

// ------------------------------------------------------------------------------------------------
// face_detection.rs
// ------------------------------------------------------------------------------------------------
struct SsdOptions { int num_layers; int input_size; int strides[4]; float interpolated_scale_aspect_ratio; };

// SSDOptions::new_front/back/short/full (face_detection.rs:39-85)
FDL_HD bool ssd_options_for(int model, SsdOptions* o) {
  switch (model) {
    case FDL_MODEL_FRONT_CAMERA: case FDL_MODEL_SHORT:
      o->num_layers = 4; o->input_size = 128; o->strides[0] = 8; o->strides[1] = o->strides[2] = o->strides[3] = 16;
      o->interpolated_scale_aspect_ratio = 1.0f; return true;
    case FDL_MODEL_BACK_CAMERA:
      o->num_layers = 4; o->input_size = 256; o->strides[0] = 16; o->strides[1] = o->strides[2] = o->strides[3] = 32;
      o->interpolated_scale_aspect_ratio = 1.0f; return true;
    case FDL_MODEL_FULL: case FDL_MODEL_FULL_SPARSE:
      o->num_layers = 1; o->input_size = 192; o->strides[0] = 4; o->strides[1] = o->strides[2] = o->strides[3] = 0;
      o->interpolated_scale_aspect_ratio = 0.0f; return true;
    default: return false;
  }
}
// Number of anchors ssd_generate_anchors (face_detection.rs:366-413) produces.
FDL_HD int ssd_num_anchors(const SsdOptions& o) {
  int n = 0, layer = 0;
  while (layer < o.num_layers) {
    int last = layer, repeats = 0;
    while (last < o.num_layers && o.strides[last] == o.strides[layer]) { ++last; repeats += o.interpolated_scale_aspect_ratio == 1.0f ? 2 : 1; }
    int fm = o.input_size / o.strides[layer];
    n += fm * fm * repeats;
    layer = last;
  }
  return n;
}
// Closed form of the same loop nest: anchor `i` -> (x_center, y_center) in f32.
FDL_HD void ssd_anchor(const SsdOptions& o, int i, float* xc, float* yc) {
  int layer = 0;
  while (layer < o.num_layers) {
    int last = layer, repeats = 0;
    while (last < o.num_layers && o.strides[last] == o.strides[layer]) { ++last; repeats += o.interpolated_scale_aspect_ratio == 1.0f ? 2 : 1; }
    int fm = o.input_size / o.strides[layer];
    int count = fm * fm * repeats;
    if (i < count) {
      int cell = i / repeats;
      int y = cell / fm, x = cell - y * fm;
      *xc = ((float)x + 0.5f) / (float)fm;
      *yc = ((float)y + 0.5f) / (float)fm;
      return;
    }
    i -= count;
    layer = last;
  }
  *xc = *yc = 0.f;
}

// decode_boxes (face_detection.rs:269-296) for one anchor: raw[16] -> data[16] ([8,2] row-major).
FDL_HD void decode_box(const float* raw, float ax, float ay, float scale, float* d) {
  for (int k = 0; k < 16; ++k) d[k] = raw[k] / scale;
  d[0] += ax; d[1] += ay;
  for (int r = 2; r < 8; ++r) { d[2 * r] += ax; d[2 * r + 1] += ay; }
  float cx = d[0], cy = d[1];
  float hx = d[2] / 2.0f, hy = d[3] / 2.0f;
  d[0] = cx - hx; d[1] = cy - hy;
  d[2] = cx + hx; d[3] = cy + hy;
}
// get_sigmoid_score (face_detection.rs:300-314) + sigmoid (transform.rs:111-113), f32.
// `(-x).exp()` is libm's expf on the reference's host (correctly rounded in practice); CUDA's expf is only
// good to 2 ulp, so the exponential is evaluated in f64 and rounded once to f32.
FDL_HD float sigmoid_f32(float x) { return 1.0f / (1.0f + (float)exp(-(double)x)); }
FDL_HD float ssd_score(float raw) {
  float x = raw < -80.0f ? -80.0f : (raw > 80.0f ? 80.0f : raw);
  return sigmoid_f32(x);
}

// ------------------------------------------------------------------------------------------------
// nms.rs:5-17 overlap_similarity on f32 boxes widened to f64 (types.rs:139-159)
// ------------------------------------------------------------------------------------------------
FDL_HD double bbox_area(double xmin, double ymin, double xmax, double ymax) {
  double w = xmax - xmin, h = ymax - ymin;
  return (w <= 0.0 || h <= 0.0) ? 0.0 : w * h;
}
FDL_HD double overlap_similarity(const float* a, const float* b) {
  double axmin = a[0], aymin = a[1], axmax = a[2], aymax = a[3];
  double bxmin = b[0], bymin = b[1], bxmax = b[2], bymax = b[3];
  double xmin = dmax(axmin, bxmin), ymin = dmax(aymin, bymin), xmax = dmin(axmax, bxmax), ymax = dmin(aymax, bymax);
  if (!(xmin < xmax && ymin < ymax)) return 0.0;
  double ia = bbox_area(xmin, ymin, xmax, ymax);
  double den = bbox_area(axmin, aymin, axmax, aymax) + bbox_area(bxmin, bymin, bxmax, bymax) - ia;
  return den > 0.0 ? ia / den : 0.0;
}

// ------------------------------------------------------------------------------------------------
// transform.rs:44-109 bbox_to_roi + select_roi_size
// ------------------------------------------------------------------------------------------------
FDL_HD bool bbox_to_roi(double xmin, double ymin, double xmax, double ymax, int img_w, int img_h, bool have_kp, double x0, double y0,
                        double x1, double y1, double scale_x, double scale_y, int size_mode, fdl_rect* out) {
  if (!(xmin >= -1.0 && xmax < 2.0 && ymin >= -1.0)) return false;   // BBox::normalized (types.rs:134-136, sic)
  // BBox::absolute (types.rs:168-173): normalized() is true here
  double axmin = xmin * (double)img_w, aymin = ymin * (double)img_h, axmax = xmax * (double)img_w, aymax = ymax * (double)img_h;
  double width = axmax - axmin, height = aymax - aymin;
  double iw = (double)img_w, ih = (double)img_h;
  if (size_mode == FDL_SIZE_MODE_SQUARE_LONG) { double l = dmax(width, height); width = l / iw; height = l / ih; }
  else if (size_mode == FDL_SIZE_MODE_SQUARE_SHORT) { double s = dmin(width, height); width = s / iw; height = s / ih; }
  width *= scale_x; height *= scale_y;
  double cx = xmin + (xmax - xmin) / 2.0, cy = ymin + (ymax - ymin) / 2.0;
  double rotation = 0.0;
  if (have_kp) {
    const double PI = 3.14159265358979323846;
    double angle = -atan2(y0 - y1, x1 - x0);
    double two_pi = 2.0 * PI;
    rotation = angle - two_pi * floor((angle + PI) / two_pi);
  }
  out->x_center = cx; out->y_center = cy; out->width = width; out->height = height; out->rotation = rotation;
  out->normalized = 1; out->_pad = 0;
  return true;
}

// face_detection_to_roi (face_landmark.rs:180-198); size_mode FDL_SIZE_MODE_NONE -> SquareLong.
FDL_HD bool face_detection_to_roi(const float* d /*[16]*/, int img_w, int img_h, int size_mode, fdl_rect* out) {
  float fw = (float)img_w, fh = (float)img_h;          // Detection::scaled_by_image_size, types.rs:237-245 (f32)
  double lx = (double)(d[4] * fw), ly = (double)(d[5] * fh), rx = (double)(d[6] * fw), ry = (double)(d[7] * fh);
  int mode = size_mode == FDL_SIZE_MODE_NONE ? FDL_SIZE_MODE_SQUARE_LONG : size_mode;
  return bbox_to_roi((double)d[0], (double)d[1], (double)d[2], (double)d[3], img_w, img_h, true, lx, ly, rx, ry, 1.5, 1.5, mode, out);
}

// one eye of iris_roi_from_face_landmarks (iris_landmark.rs:268-292): landmarks a, b (x,y as f64).
FDL_HD bool eye_roi(double ax, double ay, double bx, double by, int img_w, int img_h, fdl_rect* out) {
  double xmin = dmin(ax, bx), ymin = dmin(ay, by), xmax = dmax(ax, bx), ymax = dmax(ay, by);   // bbox_from_landmarks
  return bbox_to_roi(xmin, ymin, xmax, ymax, img_w, img_h, true, ax, ay, bx, by, 2.3, 2.3, FDL_SIZE_MODE_SQUARE_LONG, out);
}

// ------------------------------------------------------------------------------------------------
// transform.rs:351-432 project_landmarks for one (x,y,z) triple
// ------------------------------------------------------------------------------------------------
struct ProjectParams {
  float tw, th;            // tensor size as f32
  int flip;
  int unpad;               // padding != (0,0,0,0)
  double left, top, h_scale, v_scale;
  int has_roi;
  float c, s;              // cos/sin of the normalised ROI's rotation, cast to f32
  double roi_w, roi_h, roi_xc, roi_yc;
};
FDL_HD void project_setup(int tensor_w, int tensor_h, int img_w, int img_h, const double pad[4], const fdl_rect* roi, bool flip,
                          ProjectParams* p) {
  p->tw = (float)tensor_w; p->th = (float)tensor_h; p->flip = flip ? 1 : 0;
  p->unpad = !(pad[0] == 0.0 && pad[1] == 0.0 && pad[2] == 0.0 && pad[3] == 0.0);
  p->left = pad[0]; p->top = pad[1];
  p->h_scale = 1.0 - (pad[0] + pad[2]); p->v_scale = 1.0 - (pad[1] + pad[3]);
  p->has_roi = roi != nullptr;
  if (roi) {
    fdl_rect nr = rect_scaled(*roi, (double)img_w, (double)img_h, true);
    p->s = (float)sin(nr.rotation); p->c = (float)cos(nr.rotation);
    p->roi_w = nr.width; p->roi_h = nr.height; p->roi_xc = nr.x_center; p->roi_yc = nr.y_center;
  } else { p->s = 0.f; p->c = 1.f; p->roi_w = p->roi_h = 1.0; p->roi_xc = p->roi_yc = 0.0; }
}
FDL_HD void project_point(const ProjectParams& p, const float* raw, float* out) {
  float x = raw[0] / p.tw, y = raw[1] / p.th, z = raw[2] / p.tw;
  if (p.flip) x = x * -1.0f + 1.0f;
  if (p.unpad) {
    x = (float)(((double)x - p.left) / p.h_scale);
    y = (float)(((double)y - p.top) / p.v_scale);
    z = (float)(((double)z - 0.0) / p.h_scale);
  }
  if (p.has_roi) {
    float xx = x - 0.5f, yy = y - 0.5f;
    float rx = xx * p.c + yy * (-p.s);       // [x, y, 0] . [[c, s, 0], [-s, c, 0], [1, 1, 1]]  (:393-402)
    float ry = xx * p.s + yy * p.c;
    x = (float)((double)rx * p.roi_w + p.roi_xc);
    y = (float)((double)ry * p.roi_h + p.roi_yc);
    z = (float)((double)z * p.roi_w + 0.0);
  }
  out[0] = x; out[1] = y; out[2] = z;
}

// ------------------------------------------------------------------------------------------------
// iris refinement: iris_landmark.rs:64-95, :380-433
// ------------------------------------------------------------------------------------------------
// LEFT_/RIGHT_EYE_TO_FACE_LANDMARK_INDEX (iris_landmark.rs:64-95): eye contour point n of eye `eye`
// (0 = left, 1 = right) replaces this face landmark in update_face_landmarks_with_iris_results (:380-398).
FDL_HD int eye_to_face_landmark_index(int eye, int n) {
  const int16_t kLeft[FDL_NUM_EYE_CONTOUR] = {
      33,  7,   163, 144, 145, 153, 154, 155, 133, 246, 161, 160, 159, 158, 157, 173, 130, 25,  110, 24,  23,  22,  26,  112,
      243, 247, 30,  29,  27,  28,  56,  190, 226, 31,  228, 229, 230, 231, 232, 233, 244, 113, 225, 224, 223, 222, 221, 189,
      35,  124, 46,  53,  52,  65,  143, 111, 117, 118, 119, 120, 121, 128, 245, 156, 70,  63,  105, 66,  107, 55,  193};
  const int16_t kRight[FDL_NUM_EYE_CONTOUR] = {
      263, 249, 390, 373, 374, 380, 381, 382, 362, 466, 388, 387, 386, 385, 384, 398, 359, 255, 339, 254, 253, 252, 256, 341,
      463, 467, 260, 259, 257, 258, 286, 414, 446, 261, 448, 449, 450, 451, 452, 453, 464, 342, 445, 444, 443, 442, 441, 413,
      265, 353, 276, 283, 282, 295, 372, 340, 346, 347, 348, 349, 350, 357, 465, 383, 300, 293, 334, 296, 336, 285, 417};
  return eye ? kRight[n] : kLeft[n];
}
// get_iris_diameter (iris_landmark.rs:401-418): iris = 5 x (x, y, z) normalised landmarks in IrisIndex order
// (Center, Left, Top, Right, Bottom :104-110); the Landmark fields are f64 holding f32 values.
template <typename T>
FDL_HD double iris_diameter(const T* iris, int img_w, int img_h) {
  const double w = (double)img_w, h = (double)img_h;
  double dx = (double)iris[3 * 1] * w - (double)iris[3 * 3] * w, dy = (double)iris[3 * 1 + 1] * h - (double)iris[3 * 3 + 1] * h;
  const double horiz = sqrt(dx * dx + dy * dy);                 // Left - Right
  dx = (double)iris[3 * 2] * w - (double)iris[3 * 4] * w; dy = (double)iris[3 * 2 + 1] * h - (double)iris[3 * 4 + 1] * h;
  const double vert = sqrt(dx * dx + dy * dy);                  // Top - Bottom
  return (vert + horiz) / 2.0;
}
// get_iris_depth (iris_landmark.rs:421-433); the image centre is (width / 2, height / 2) in INTEGER arithmetic (:426).
template <typename T>
FDL_HD double iris_depth(const T* iris, double focal_length_mm, double iris_size_px, int img_w, int img_h) {
  const double x0 = (double)(img_w / 2), y0 = (double)(img_h / 2);
  const double x1 = (double)iris[0] * (double)img_w, y1 = (double)iris[1] * (double)img_h;
  const double y = sqrt((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1));
  const double x = sqrt(focal_length_mm * focal_length_mm + y * y);
  return 11.8 * x / iris_size_px;                               // IRIS_SIZE_IN_MM :100
}

}  // namespace fdl
