// jpeg_color.cuh -- the pixel arithmetic of the JPEG colour stage (jdsample.c fancy upsampling + jdcolor.c YCbCr -> RGB) for four
// pixels of one row, shared by the colour kernels of jpeg_kernels.cu (whole frames, listed rows, ROI row spans).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "jpeg_device.h"

namespace fdl {
namespace {

// One image of the colour fast path (both chroma components at 2x2 or at 2x1, rows 16-byte aligned): plane / output pointers, strides.
struct ColorShared {
  const uint8_t* py; const uint8_t* pcb; const uint8_t* pcr;
  uint8_t* out;
  int sy, sc, width, height, out_stride, cw, ch, v2;
};

// bytes (c-1, c, c+1, c+2) of a plane row as one word (byte 0 = c-1); c even.  For c == 0 byte 0 is unspecified (the edge rule never uses it).
__device__ __forceinline__ uint32_t row4(const uint8_t* __restrict__ row, int c) {
  if (c == 0) return __ldg(reinterpret_cast<const uint32_t*>(row)) << 8;
  const int a = c - 1;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(row + (a & ~3));
  return __funnelshift_r(__ldg(p), __ldg(p + 1), (a & 3) * 8);
}
// jdcolor.c ycc_rgb_convert for four pixels -> 12 bytes r g b r g b ... in three words.  The operations are regrouped, not changed:
// y + ((k * x + ONE_HALF) >> 16) == ((y << 16) + ONE_HALF + k * x) >> 16 (arithmetic shift = floor), `yh` is (y << 16) + ONE_HALF
// built by one byte permute, x = Cb - 128 / Cr - 128 arrive already centred, the clamp to 0..255 is the saturating pack.
__device__ __forceinline__ uint32_t pack_sat_u8(int hi, int lo) {       // sat_u8(lo) | sat_u8(hi) << 8
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(0));
  return d;
}
template <int kByte>
__device__ __forceinline__ void ycc_px(uint32_t yw, int xb, int xr, int& r, int& g, int& b) {
  const int yh = (int)__byte_perm(yw, 0x00008000u, 0x6054 | (kByte << 8));
  r = (91881 * xr + yh) >> 16;
  g = (-22554 * xb + (-46802) * xr + yh) >> 16;
  b = (116130 * xb + yh) >> 16;
}
__device__ __forceinline__ void ycc_px4(uint32_t yw, const int* xb, const int* xr, uint32_t* out) {
  int r0, g0, b0, r1, g1, b1, r2, g2, b2, r3, g3, b3;
  ycc_px<0>(yw, xb[0], xr[0], r0, g0, b0);
  ycc_px<1>(yw, xb[1], xr[1], r1, g1, b1);
  ycc_px<2>(yw, xb[2], xr[2], r2, g2, b2);
  ycc_px<3>(yw, xb[3], xr[3], r3, g3, b3);
  out[0] = __byte_perm(pack_sat_u8(g0, r0), pack_sat_u8(r1, b0), 0x5410);
  out[1] = __byte_perm(pack_sat_u8(b1, g1), pack_sat_u8(g2, r2), 0x5410);
  out[2] = __byte_perm(pack_sat_u8(r3, b2), pack_sat_u8(b3, g3), 0x5410);
}


__device__ __forceinline__ void color_shared_fill(ColorShared& P, const JpegImageDesc& d, const uint8_t* planes, uint8_t* out) {
  P.py = planes + d.plane_off[0]; P.pcb = planes + d.plane_off[1]; P.pcr = planes + d.plane_off[2];
  P.out = out + d.out_off;
  P.sy = d.bcols[0] * 8; P.sc = d.bcols[1] * 8; P.width = d.width; P.height = d.height; P.out_stride = d.out_stride;
  P.cw = d.cw[1]; P.ch = d.ch[1]; P.v2 = d.vmax / d.vs[1] == 2;
}


// Four pixels x0 .. x0 + 3 (x0 a multiple of 4) of one row: `yrow` the luma row, `nb` / `nr` the near chroma rows (y >> 1, or y for
// h2v1), `fb` / `fr` the far ones (y >> 1 - 1 for even y, + 1 for odd y, clamped; unused for h2v1) -> r g b r g b ... in three words.
__device__ __forceinline__ void color_px4(const ColorShared& P, const uint8_t* __restrict__ yrow, const uint8_t* __restrict__ nb,
                                          const uint8_t* __restrict__ nr, const uint8_t* __restrict__ fb, const uint8_t* __restrict__ fr,
                                          int x0, uint32_t* rgb) {
  const int cx = x0 >> 1, cw = P.cw;
  // jdsample.c's fancy upsampling as dot products over the row words (bytes: samples c-1, c, c+1, c+2).  Far-row weights of the
  // four outputs; the near row takes three times them (h2v2) or is the only row (h2v1).  The image edges replace the missing
  // neighbour by the sample itself: weight 4 on it.  The rounding constant carries -128 (scaled), so the shift yields Cb - 128.
  const uint32_t k0 = cx == 0 ? 0x00000400u : 0x00000301u, k1 = cx == cw - 1 ? 0x00000400u : 0x00010300u, k2 = 0x00030100u,
                 k3 = cx + 1 >= cw - 1 ? 0x00040000u : 0x01030000u;
  const uint32_t yw = __ldg(reinterpret_cast<const uint32_t*>(yrow + x0));
  int xbv[4], xrv[4];
  if (P.v2) {
    const uint32_t b8 = (uint32_t)(8 - 128 * 16), b7 = (uint32_t)(7 - 128 * 16);
    const uint32_t wnb = row4(nb, cx), wfb = row4(fb, cx), wnr = row4(nr, cx), wfr = row4(fr, cx);
    xbv[0] = (int)__dp4a(wfb, k0, __dp4a(wnb, 3u * k0, b8)) >> 4; xbv[1] = (int)__dp4a(wfb, k1, __dp4a(wnb, 3u * k1, b7)) >> 4;
    xbv[2] = (int)__dp4a(wfb, k2, __dp4a(wnb, 3u * k2, b8)) >> 4; xbv[3] = (int)__dp4a(wfb, k3, __dp4a(wnb, 3u * k3, b7)) >> 4;
    xrv[0] = (int)__dp4a(wfr, k0, __dp4a(wnr, 3u * k0, b8)) >> 4; xrv[1] = (int)__dp4a(wfr, k1, __dp4a(wnr, 3u * k1, b7)) >> 4;
    xrv[2] = (int)__dp4a(wfr, k2, __dp4a(wnr, 3u * k2, b8)) >> 4; xrv[3] = (int)__dp4a(wfr, k3, __dp4a(wnr, 3u * k3, b7)) >> 4;
  } else {
    const uint32_t b1 = (uint32_t)(1 - 128 * 4), b2 = (uint32_t)(2 - 128 * 4);
    const uint32_t wnb = row4(nb, cx), wnr = row4(nr, cx);
    xbv[0] = (int)__dp4a(wnb, k0, b1) >> 2; xbv[1] = (int)__dp4a(wnb, k1, b2) >> 2;
    xbv[2] = (int)__dp4a(wnb, k2, b1) >> 2; xbv[3] = (int)__dp4a(wnb, k3, b2) >> 2;
    xrv[0] = (int)__dp4a(wnr, k0, b1) >> 2; xrv[1] = (int)__dp4a(wnr, k1, b2) >> 2;
    xrv[2] = (int)__dp4a(wnr, k2, b1) >> 2; xrv[3] = (int)__dp4a(wnr, k3, b2) >> 2;
  }
  ycc_px4(yw, xbv, xrv, rgb);
}

// the chroma rows of luma row y
__device__ __forceinline__ void color_row_ptrs(const ColorShared& P, int y, const uint8_t** yrow, const uint8_t** nb, const uint8_t** nr,
                                               const uint8_t** fb, const uint8_t** fr) {
  const bool v2 = P.v2 != 0;
  const int cy = v2 ? y >> 1 : y;
  const int fy = v2 ? ((y & 1) ? min(cy + 1, P.ch - 1) : max(cy - 1, 0)) : cy;
  *yrow = P.py + (long long)y * P.sy;
  *nb = P.pcb + (long long)cy * P.sc; *nr = P.pcr + (long long)cy * P.sc;
  *fb = P.pcb + (long long)fy * P.sc; *fr = P.pcr + (long long)fy * P.sc;
}

}  // namespace
}  // namespace fdl
