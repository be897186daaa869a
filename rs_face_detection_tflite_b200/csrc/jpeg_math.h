// jpeg_math.h -- the per-block / per-pixel arithmetic of baseline JPEG decoding after the entropy stage, written once for host
// and device (the glue_math.h pattern): dequantisation + libjpeg's accurate integer inverse DCT, "fancy" chroma upsampling in
// gather form (one output sample from its <= 4 source samples), fixed-point YCbCr -> RGB.
//
// This is the back half of the frame-ingest row: the reference decodes frames with
// imgcodecs::imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB) (src/face_detection_lite/utils.rs:8-21), i.e. OpenCV's bundled libjpeg
// with its default choices (JDCT_ISLOW, do_fancy_upsampling).  Bit-exact with cv2.imdecode: tests/test_oracle_jpeg.py drives these
// functions on the host through tests/hostcheck.  NOT yet wired into libfdl_b200.so: no kernel calls it, there is no device
// entropy decoder -- frames still enter the library decoded.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FDL_JHD __host__ __device__ __forceinline__
#else
#define FDL_JHD inline
#endif

namespace fdl {

// jidctint.c (CONST_BITS 13, PASS1_BITS 2): one 1-D pass over 8 values spaced `stride` apart, in place on int32.
FDL_JHD void jpeg_idct_1d(int* d, int stride, int in_shift, int descale) {
  const int i0 = d[0], i1 = d[stride], i2 = d[2 * stride], i3 = d[3 * stride], i4 = d[4 * stride], i5 = d[5 * stride], i6 = d[6 * stride],
            i7 = d[7 * stride];
  // even part
  int z1 = (i2 + i6) * 4433;                        // FIX_0_541196100
  const int tmp2 = z1 + i6 * (-15137);              // FIX_1_847759065
  const int tmp3 = z1 + i2 * 6270;                  // FIX_0_765366865
  const int tmp0 = (int)((unsigned)(i0 + i4) << in_shift), tmp1 = (int)((unsigned)(i0 - i4) << in_shift);
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  // odd part
  int t0 = i7, t1 = i5, t2 = i3, t3 = i1;
  z1 = t0 + t3;
  int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const int z5 = (z3 + z4) * 9633;                  // FIX_1_175875602
  t0 *= 2446; t1 *= 16819; t2 *= 25172; t3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  const int half = 1 << (descale - 1);
  d[0] = (tmp10 + t3 + half) >> descale;            d[7 * stride] = (tmp10 - t3 + half) >> descale;
  d[stride] = (tmp11 + t2 + half) >> descale;       d[6 * stride] = (tmp11 - t2 + half) >> descale;
  d[2 * stride] = (tmp12 + t1 + half) >> descale;   d[5 * stride] = (tmp12 - t1 + half) >> descale;
  d[3 * stride] = (tmp13 + t0 + half) >> descale;   d[4 * stride] = (tmp13 - t0 + half) >> descale;
}

// range_limit[(v) & RANGE_MASK] of jdmaster.c's table, centred on +128
FDL_JHD uint8_t jpeg_range_limit(int v) {
  v &= 0x3FF;
  return (uint8_t)(v < 128 ? v + 128 : (v < 512 ? 255 : (v < 896 ? 0 : v - 896)));
}

// jpeg_idct_islow: quantised coefficients (natural order) x quantisation table (natural order) -> 8x8 samples
FDL_JHD void jpeg_idct_islow_8x8(const int16_t* coef, const uint16_t* quant, uint8_t* out, int out_stride) {
  int ws[64];
  for (int i = 0; i < 64; ++i) ws[i] = (int)coef[i] * (int)quant[i];
  for (int c = 0; c < 8; ++c) jpeg_idct_1d(ws + c, 8, 13, 13 - 2);               // columns
  for (int r = 0; r < 8; ++r) {
    jpeg_idct_1d(ws + 8 * r, 1, 13, 13 + 2 + 3);                                  // rows
    for (int c = 0; c < 8; ++c) out[r * out_stride + c] = jpeg_range_limit(ws[8 * r + c]);
  }
}

// jdsample.c h2v2_fancy_upsample, gather form: the full-resolution sample (x, y) of a component stored at half resolution in
// both directions.  cw x ch is the REAL downsampled size (ceil(image / 2)), whose edges are the ones replicated.
// (jinit_upsampler picks the fancy method only when the downsampled component is more than 2 samples wide: plain replication else.)
FDL_JHD int jpeg_h2v2_fancy_at(const uint8_t* plane, int stride, int cw, int ch, int x, int y) {
  const int cy = y >> 1, cx = x >> 1;
  if (cw <= 2) return plane[(long long)cy * stride + cx];
  int fy = (y & 1) ? cy + 1 : cy - 1;
  fy = fy < 0 ? 0 : (fy > ch - 1 ? ch - 1 : fy);
  const uint8_t* near_row = plane + (long long)cy * stride;
  const uint8_t* far_row = plane + (long long)fy * stride;
  const int s = 3 * near_row[cx] + far_row[cx];
  if (!(x & 1)) {
    if (cx == 0) return (4 * s + 8) >> 4;
    return (3 * s + (3 * near_row[cx - 1] + far_row[cx - 1]) + 8) >> 4;
  }
  if (cx == cw - 1) return (4 * s + 7) >> 4;
  return (3 * s + (3 * near_row[cx + 1] + far_row[cx + 1]) + 7) >> 4;
}

// h2v1_fancy_upsample, gather form (4:2:2)
FDL_JHD int jpeg_h2v1_fancy_at(const uint8_t* plane, int stride, int cw, int x, int y) {
  const uint8_t* row = plane + (long long)y * stride;
  const int cx = x >> 1, s = row[cx];
  if (cw <= 2) return s;
  if (!(x & 1)) return cx == 0 ? s : (3 * s + row[cx - 1] + 1) >> 2;
  return cx == cw - 1 ? s : (3 * s + row[cx + 1] + 2) >> 2;
}

// jdcolor.c ycc_rgb_convert (SCALEBITS 16; the tables evaluated in place; >> is arithmetic, as libjpeg's RIGHT_SHIFT)
FDL_JHD void jpeg_ycc_to_rgb(int y, int cb, int cr, uint8_t* rgb) {
  const int xb = cb - 128, xr = cr - 128;
  const int r = y + ((91881 * xr + 32768) >> 16);                       // FIX(1.40200)
  const int g = y + ((-22554 * xb + 32768 + (-46802) * xr) >> 16);      // FIX(0.34414), FIX(0.71414)
  const int b = y + ((116130 * xb + 32768) >> 16);                      // FIX(1.77200)
  rgb[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  rgb[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
  rgb[2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
}

}  // namespace fdl
