// jpeg_parse.h -- host-side marker parser for the frame-ingest row: what a decoder needs to know before
// the entropy stage.  Accepts exactly what oracle/jpeg_decode.py accepts -- 8-bit baseline (SOF0 / SOF1 Huffman) files with one
// interleaved scan, 1 or 3 components, luma at full resolution and chroma at 1x1, 2x1 or 2x2, EXIF orientation 1 or none -- and
// rejects the rest with a message (imdecode would rotate EXIF-oriented files; progressive files need another entropy stage).
// Header-only and CUDA-free so that tests/hostcheck can compile it; NOT yet used by libfdl_b200.so.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

#include "jpeg_math.h"

namespace fdl {

struct JpegComponent {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
};

struct JpegHeader {
  int width = 0, height = 0, ncomp = 0;
  JpegComponent comp[3];
  int hmax = 1, vmax = 1;
  int mcus_x = 0, mcus_y = 0;
  int restart_interval = 0;
  uint16_t quant[4][64];        // natural (row-major) order
  bool have_quant[4] = {false, false, false, false};
  uint8_t dht[2][4][16 + 256];  // [class][id]: counts[16] + symbols
  bool have_dht[2][4] = {{false, false, false, false}, {false, false, false, false}};
  size_t scan_offset = 0;       // first entropy-coded byte
};

inline int jpeg_exif_orientation(const uint8_t* t, size_t n) {
  if (n < 8) return 0;
  const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
  if (!le && !be) return 0;
  auto u16 = [&](size_t o) { return le ? (unsigned)(t[o] | (t[o + 1] << 8)) : (unsigned)((t[o] << 8) | t[o + 1]); };
  auto u32 = [&](size_t o) { return le ? (u16(o) | (u16(o + 2) << 16)) : ((u16(o) << 16) | u16(o + 2)); };
  const size_t off = u32(4);
  if (off + 2 > n) return 0;
  const unsigned cnt = u16(off);
  for (unsigned k = 0; k < cnt; ++k) {
    const size_t e = off + 2 + 12 * (size_t)k;
    if (e + 12 > n) break;
    if (u16(e) == 0x0112 && u16(e + 2) == 3) return (int)u16(e + 8);
  }
  return 0;
}

// Returns true and fills `h`, or false with a reason in `err`.
inline bool jpeg_parse_header(const uint8_t* d, size_t n, JpegHeader* h, std::string* err) {
  static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  auto fail = [&](const char* m) { if (err) *err = m; return false; };
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG");
  size_t p = 2;
  bool have_frame = false;
  while (p + 4 <= n) {
    if (d[p] != 0xFF) return fail("marker expected");
    const int m = d[p + 1];
    if (m == 0xFF) { ++p; continue; }
    const size_t L = ((size_t)d[p + 2] << 8) | d[p + 3];
    if (L < 2 || p + 2 + L > n) return fail("truncated segment");
    const uint8_t* s = d + p + 4;
    const size_t sl = L - 2;
    if (m == 0xDB) {
      for (size_t q = 0; q < sl;) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        if (tq > 3 || q + 1 + (pq ? 128u : 64u) > sl) return fail("bad DQT");
        for (int i = 0; i < 64; ++i) h->quant[tq][zz[i]] = pq ? (uint16_t)((s[q + 1 + 2 * i] << 8) | s[q + 2 + 2 * i]) : s[q + 1 + i];
        h->have_quant[tq] = true;
        q += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || s[0] != 8) return fail("only 8-bit samples");
      h->height = (s[1] << 8) | s[2]; h->width = (s[3] << 8) | s[4]; h->ncomp = s[5];
      if (h->ncomp != 1 && h->ncomp != 3) return fail("1 or 3 components only");
      if (sl < 6 + 3 * (size_t)h->ncomp || h->width == 0 || h->height == 0) return fail("bad SOF");
      for (int c = 0; c < h->ncomp; ++c) {
        h->comp[c].id = s[6 + 3 * c]; h->comp[c].h = s[7 + 3 * c] >> 4; h->comp[c].v = s[7 + 3 * c] & 15; h->comp[c].tq = s[8 + 3 * c];
        if (h->comp[c].tq > 3 || h->comp[c].h < 1 || h->comp[c].v < 1) return fail("bad SOF");
      }
      have_frame = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return fail("only baseline sequential Huffman JPEG (SOF0/SOF1)");
    } else if (m == 0xC4) {
      for (size_t q = 0; q < sl;) {
        const int tc = s[q] >> 4, th = s[q] & 15;
        if (tc > 1 || th > 3 || q + 17 > sl) return fail("bad DHT");
        size_t nsym = 0;
        for (int i = 0; i < 16; ++i) nsym += s[q + 1 + i];
        if (nsym > 256 || q + 17 + nsym > sl) return fail("bad DHT");
        if (!jpeg_huff_valid(s + q + 1)) return fail("bad DHT: code lengths over-subscribed");
        std::memset(h->dht[tc][th], 0, sizeof(h->dht[tc][th]));
        std::memcpy(h->dht[tc][th], s + q + 1, 16 + nsym);
        h->have_dht[tc][th] = true;
        q += 17 + nsym;
      }
    } else if (m == 0xDD) {
      if (sl < 2) return fail("bad DRI");
      h->restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xE1 && sl > 6 && std::memcmp(s, "Exif\0\0", 6) == 0) {
      const int o = jpeg_exif_orientation(s + 6, sl - 6);
      if (o != 0 && o != 1) return fail("EXIF orientation: not restated");
    } else if (m == 0xDA) {
      if (!have_frame) return fail("SOS before SOF");
      if (sl < 1 || s[0] != h->ncomp || sl < 1 + 2 * (size_t)h->ncomp) return fail("only single-scan (interleaved) files");
      for (int c = 0; c < h->ncomp; ++c) {
        if (s[1 + 2 * c] != h->comp[c].id) return fail("scan component order");
        h->comp[c].td = s[2 + 2 * c] >> 4; h->comp[c].ta = s[2 + 2 * c] & 15;
        if (h->comp[c].td > 3 || h->comp[c].ta > 3 || !h->have_dht[0][h->comp[c].td] || !h->have_dht[1][h->comp[c].ta]) return fail("missing Huffman table");
        if (!h->have_quant[h->comp[c].tq]) return fail("missing quantisation table");
      }
      h->scan_offset = p + 2 + L;
      h->hmax = h->vmax = 1;
      for (int c = 0; c < h->ncomp; ++c) { if (h->comp[c].h > h->hmax) h->hmax = h->comp[c].h; if (h->comp[c].v > h->vmax) h->vmax = h->comp[c].v; }
      for (int c = 0; c < h->ncomp; ++c) {
        const int eh = h->hmax / h->comp[c].h, ev = h->vmax / h->comp[c].v;
        if (h->hmax % h->comp[c].h || h->vmax % h->comp[c].v || !((eh == 1 && ev == 1) || (eh == 2 && ev == 1) || (eh == 2 && ev == 2)))
          return fail("unsupported sampling factors");
      }
      h->mcus_x = (h->width + 8 * h->hmax - 1) / (8 * h->hmax);
      h->mcus_y = (h->height + 8 * h->vmax - 1) / (8 * h->vmax);
      return true;
    }
    p += 2 + L;
  }
  return fail("no frame / scan");
}

}  // namespace fdl
