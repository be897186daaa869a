// jpeg_math.h -- the per-block / per-pixel arithmetic of baseline JPEG decoding after the entropy stage, written once for host
// and device (the glue_math.h pattern): dequantisation + libjpeg's accurate integer inverse DCT, "fancy" chroma upsampling in
// gather form (one output sample from its <= 4 source samples), fixed-point YCbCr -> RGB.
//
// This is the back half of the frame-ingest row: the reference decodes frames with
// imgcodecs::imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB) (src/face_detection_lite/utils.rs:8-21), i.e. OpenCV's bundled libjpeg
// with its default choices (JDCT_ISLOW, do_fancy_upsampling).  Bit-exact with cv2.imdecode: tests/test_oracle_jpeg.py drives these
// functions on the host through tests/hostcheck.  NOT yet wired into libfdl_b200.so: no kernel calls it -- frames still enter
// the library decoded.  The entropy-stage building blocks at the end (table, bit reader, one-block decoder) are the sequential
// core a device decoder runs per restart interval / subsequence; they are checked on the host the same way.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FDL_JHD __host__ __device__ __forceinline__
#else
#define FDL_JHD inline
#endif

namespace fdl {

// jidctint.c (CONST_BITS 13, PASS1_BITS 2): one 1-D pass over 8 values spaced `stride` apart, in place on int32.
// All sums and products are taken modulo 2^32 (unsigned): for the coefficients of a valid file nothing overflows and the bits are
// libjpeg's; for corrupt data (coefficient x quantiser products up to 2^27) the result is garbage but defined -- no signed overflow.
FDL_JHD void jpeg_idct_1d(int* d, int stride, int in_shift, int descale) {
  typedef unsigned U;
  const U i0 = (U)d[0], i1 = (U)d[stride], i2 = (U)d[2 * stride], i3 = (U)d[3 * stride], i4 = (U)d[4 * stride], i5 = (U)d[5 * stride],
          i6 = (U)d[6 * stride], i7 = (U)d[7 * stride];
  // even part
  U z1 = (i2 + i6) * 4433u;                         // FIX_0_541196100
  const U tmp2 = z1 + i6 * (U)(-15137);             // FIX_1_847759065
  const U tmp3 = z1 + i2 * 6270u;                   // FIX_0_765366865
  const U tmp0 = (i0 + i4) << in_shift, tmp1 = (i0 - i4) << in_shift;
  const U tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  // odd part
  U t0 = i7, t1 = i5, t2 = i3, t3 = i1;
  z1 = t0 + t3;
  U z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const U z5 = (z3 + z4) * 9633u;                   // FIX_1_175875602
  t0 *= 2446u; t1 *= 16819u; t2 *= 25172u; t3 *= 12299u;
  z1 *= (U)(-7373); z2 *= (U)(-20995); z3 = z3 * (U)(-16069) + z5; z4 = z4 * (U)(-3196) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  const U half = 1u << (descale - 1);
  d[0] = (int)(tmp10 + t3 + half) >> descale;            d[7 * stride] = (int)(tmp10 - t3 + half) >> descale;
  d[stride] = (int)(tmp11 + t2 + half) >> descale;       d[6 * stride] = (int)(tmp11 - t2 + half) >> descale;
  d[2 * stride] = (int)(tmp12 + t1 + half) >> descale;   d[5 * stride] = (int)(tmp12 - t1 + half) >> descale;
  d[3 * stride] = (int)(tmp13 + t0 + half) >> descale;   d[4 * stride] = (int)(tmp13 - t0 + half) >> descale;
}

// range_limit[(v) & RANGE_MASK] of jdmaster.c's table, centred on +128
FDL_JHD uint8_t jpeg_range_limit(int v) {
  v &= 0x3FF;
  return (uint8_t)(v < 128 ? v + 128 : (v < 512 ? 255 : (v < 896 ? 0 : v - 896)));
}

// jpeg_idct_islow: quantised coefficients (natural order) x quantisation table (natural order) -> 8x8 samples
FDL_JHD void jpeg_idct_islow_8x8(const int16_t* coef, const uint16_t* quant, uint8_t* out, int out_stride) {
  int ws[64];
  for (int i = 0; i < 64; ++i) ws[i] = (int)((unsigned)(int)coef[i] * (unsigned)quant[i]);
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

// ---- entropy stage building blocks (T.81 F.2.2; sequential within one restart interval / subsequence) ----------------------
// Huffman table in the two-level form libjpeg's decoder uses: a 9-bit lookahead table answers the common short codes in one
// probe, longer codes walk maxcode[] (T.81 F.2.2.3 / Annex C).  600 bytes per table: four of them fit shared memory many times over.
struct JpegHuff {
  uint16_t look[512];      // (length << 8) | symbol for codes of <= 9 bits (by their 9-bit prefix), 0 otherwise
  int32_t maxcode[18];     // largest code of each length (-1: none); [17] = sentinel
  int32_t valoffset[17];   // huffval index of the first code of each length minus that code
  uint32_t limit[17];      // limit[l] = the first unused code of length l, left-aligned to 16 bits: a 16-bit peek v holds a code of
                           // at most l bits iff v < limit[l] (non-decreasing in l; 0x10000 when the code space is full)
  uint8_t huffval[256];
};

// A DHT segment is usable when its code lengths describe a prefix code (no length over-subscribed: T.81 Annex C generates
// codes by counting, so `code` must stay below 2^l at every length) with at most 256 symbols.  libjpeg's jpeg_make_d_derived_tbl
// rejects the same tables ("Bogus Huffman table definition").
FDL_JHD bool jpeg_huff_valid(const uint8_t* counts) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    code += counts[l - 1];
    k += counts[l - 1];
    if (code > (1 << l)) return false;
    code <<= 1;
  }
  return k <= 256;
}

// counts[16] / symbols as in a DHT segment (callers check jpeg_huff_valid first; an invalid table is still built without
// writing out of bounds)
FDL_JHD void jpeg_huff_build(const uint8_t* counts, const uint8_t* symbols, JpegHuff* t) {
  for (int i = 0; i < 512; ++i) t->look[i] = 0;
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t->valoffset[l] = k - code;
    for (int i = 0; i < counts[l - 1] && k < 256; ++i, ++k, ++code) {
      t->huffval[k] = symbols[k];
      if (l <= 9) {
        const int lo = code << (9 - l);
        for (int j = 0; j < (1 << (9 - l)) && lo + j < 512; ++j) t->look[lo + j] = (uint16_t)((l << 8) | symbols[k]);
      }
    }
    t->maxcode[l] = counts[l - 1] ? code - 1 : -1;
    t->limit[l] = (uint32_t)code << (16 - l);
    code <<= 1;
  }
  t->maxcode[17] = 0x7FFFFFFF;
  t->maxcode[0] = -1; t->valoffset[0] = 0; t->limit[0] = 0;
  for (int i = k; i < 256; ++i) t->huffval[i] = 0;
}

// MSB-first bit reader over entropy-coded bytes: 0xFF00 -> 0xFF, any other marker stops the stream (zero bits from there on).
struct JpegBits {
  const uint8_t* p; const uint8_t* end;
  uint64_t acc; int n;
};
FDL_JHD void jpeg_bits_init(JpegBits* b, const uint8_t* p, const uint8_t* end) { b->p = p; b->end = end; b->acc = 0; b->n = 0; }
FDL_JHD void jpeg_bits_fill(JpegBits* b) {
  while (b->n <= 48) {
    unsigned v = 0;
    if (b->p < b->end) {
      v = *b->p;
      if (v == 0xFF) {
        const unsigned nx = b->p + 1 < b->end ? b->p[1] : 0xD9u;
        if (nx == 0) b->p += 2; else v = 0;          // a marker: feed zeros, stay on it
      } else {
        ++b->p;
      }
    }
    b->acc = (b->acc << 8) | v;
    b->n += 8;
  }
}
FDL_JHD int jpeg_bits_peek(JpegBits* b, int k) { if (b->n < k) jpeg_bits_fill(b); return (int)((b->acc >> (b->n - k)) & ((1u << k) - 1)); }
FDL_JHD int jpeg_bits_get(JpegBits* b, int k) { if (k == 0) return 0; const int v = jpeg_bits_peek(b, k); b->n -= k; return v; }
// RSTn: drop the remaining bits and step over the marker (T.81 E.2.4)
FDL_JHD void jpeg_bits_restart(JpegBits* b) {
  b->acc = 0; b->n = 0;
  while (b->p + 1 < b->end && !(b->p[0] == 0xFF && b->p[1] >= 0xD0 && b->p[1] <= 0xD7)) ++b->p;
  b->p += 2;
}

FDL_JHD int jpeg_huff_decode(JpegBits* b, const JpegHuff& t) {
  const int look = t.look[jpeg_bits_peek(b, 9)];
  if (look) { b->n -= look >> 8; return look & 255; }
  int code = jpeg_bits_get(b, 10), l = 10;
  while (code > t.maxcode[l]) { code = (code << 1) | jpeg_bits_get(b, 1); ++l; }
  if (l > 16) return 0;                               // corrupt data: libjpeg substitutes a zero symbol
  return t.huffval[(code + t.valoffset[l]) & 255];
}
FDL_JHD int jpeg_extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

// One 8x8 block: coef[64] in natural order, zeroed by the caller; *pred is the component's running DC predictor.
FDL_JHD void jpeg_decode_block(JpegBits* b, const JpegHuff& dc, const JpegHuff& ac, int* pred, int16_t* coef) {
  const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  int t = jpeg_huff_decode(b, dc);
  if (t > 16) t = 16;                                  // corrupt table symbol: the same clamp as jpeg_sync_step (no over-wide shifts)
  if (t) *pred += jpeg_extend(jpeg_bits_get(b, t), t);
  coef[0] = (int16_t)*pred;
  for (int k = 1; k < 64;) {
    const int rs = jpeg_huff_decode(b, ac), r = rs >> 4, s = rs & 15;
    if (s == 0) {
      if (r != 15) break;                             // EOB
      k += 16;                                        // ZRL
      continue;
    }
    k += r;
    if (k > 63) break;
    coef[zz[k]] = (int16_t)jpeg_extend(jpeg_bits_get(b, s), s);
    ++k;
  }
}

// ---- self-synchronising parallel entropy decoding (files without restart markers) ----------------------------------------
// Huffman streams resynchronise: a decoder started at an arbitrary bit with a guessed state soon falls into step with the true
// symbol sequence.  The scan (byte-unstuffed into a clean buffer first) is cut into fixed windows; every window is decoded from
// a guessed entry state, the exit states are handed to the next windows, and windows whose entry changed are decoded again
// until nothing changes (Weissenberger & Schmidt, "Massively Parallel Huffman Decoding on GPUs", ICPP 2018, and its JPEG
// follow-up).  These are the per-window primitives; tests/hostcheck runs the schedule on the host.
struct JpegSyncState {
  long long pos;   // bit position in the unstuffed scan
  int b;           // block index inside the MCU (scan order)
  int k;           // next zig-zag index of that block (0: the DC symbol is next)
};
FDL_JHD bool operator==(const JpegSyncState& x, const JpegSyncState& y) { return x.pos == y.pos && x.b == y.b && x.k == y.k; }

FDL_JHD unsigned jpeg_peek_clean(const uint8_t* buf, long long nbits, long long pos, int k) {   // k <= 24; zeros past the end
  unsigned long long w = 0;
  const long long byte = pos >> 3;
  for (int i = 0; i < 5; ++i) { const long long j = byte + i; w = (w << 8) | (unsigned long long)((j << 3) < nbits ? buf[j] : 0); }
  return (unsigned)((w >> (40 - (int)(pos & 7) - k)) & ((1u << k) - 1));
}

// One symbol (Huffman code + its extra bits) at st->pos for the block / zig-zag position in st.  Advances st; reports what the
// symbol meant: *zz = zig-zag index of the coefficient it carries (-1: none, i.e. EOB / ZRL / zero DC difference), *value = the
// coefficient (AC) or DC difference, *block_done = the symbol ended its block.
FDL_JHD void jpeg_sync_step(const uint8_t* buf, long long nbits, const JpegHuff* tabs /*[2*comp + {0 dc, 1 ac}]*/, const uint8_t* comp_of_block,
                            int blocks_per_mcu, JpegSyncState* st, int* zz, int* value, bool* block_done) {
  const JpegHuff& t = tabs[2 * comp_of_block[st->b] + (st->k ? 1 : 0)];
  int sym, len;
  const int look = t.look[jpeg_peek_clean(buf, nbits, st->pos, 9)];
  if (look) { len = look >> 8; sym = look & 255; }
  else {
    const unsigned w = jpeg_peek_clean(buf, nbits, st->pos, 16);
    len = 10;
    int code = (int)(w >> 6);
    while (len <= 16 && code > t.maxcode[len]) { ++len; if (len <= 16) code = (int)(w >> (16 - len)); }
    if (len > 16) { len = 16; sym = 0; } else sym = t.huffval[(code + t.valoffset[len]) & 255];
  }
  st->pos += len;
  *zz = -1; *value = 0; *block_done = false;
  if (st->k == 0) {                                   // DC: sym = number of extra bits
    if (sym) { const int sz = sym > 16 ? 16 : sym; *value = jpeg_extend((int)jpeg_peek_clean(buf, nbits, st->pos, sz), sz); st->pos += sz; }
    *zz = 0;
    st->k = 1;
    return;
  }
  const int r = sym >> 4, s2 = sym & 15;
  if (s2 == 0) {
    if (r != 15) *block_done = true;                  // EOB
    else { st->k += 16; if (st->k > 63) *block_done = true; }
  } else {
    st->k += r;
    if (st->k > 63) *block_done = true;               // corrupt run: the block ends (as the sequential decoder does)
    else {
      *zz = st->k;
      *value = jpeg_extend((int)jpeg_peek_clean(buf, nbits, st->pos, s2), s2);
      st->pos += s2;
      if (++st->k > 63) *block_done = true;
    }
  }
  if (*block_done) { st->k = 0; st->b = st->b + 1 == blocks_per_mcu ? 0 : st->b + 1; }
}

}  // namespace fdl
