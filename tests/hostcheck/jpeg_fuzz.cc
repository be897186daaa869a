// jpeg_fuzz.cc -- negative / fuzz test of the JPEG host headers (csrc/jpeg_parse.h, csrc/jpeg_math.h) under AddressSanitizer and
// UBSan: malformed DHT segments (over-subscribed code lengths, too many symbols), corrupt Huffman symbols (DC categories up to 255),
// damaged SOF / DQT / SOS segments and damaged entropy-coded data must be rejected or decoded to garbage -- never read or written
// out of bounds, never shifted by a negative or over-wide amount.  Test infrastructure only.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=all jpeg_fuzz.cc -o jpeg_fuzz && ./jpeg_fuzz file.jpg...
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../rs_face_detection_tflite_b200/csrc/jpeg_parse.h"

using namespace fdl;

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

// Whole decode through the headers, the way hostcheck does it, bounded by the block count.
static int decode(const std::vector<uint8_t>& f, long long* accepted) {
  JpegHeader hd;
  std::string msg;
  if (!jpeg_parse_header(f.data(), f.size(), &hd, &msg)) return 0;
  if ((long long)hd.mcus_x * hd.mcus_y > 20000) return 0;                  // a damaged SOF may claim a huge image: not this test's business
  ++*accepted;
  JpegHuff tabs[6];
  uint8_t comp_of_block[16];
  int bpm = 0;
  for (int c = 0; c < hd.ncomp; ++c) {
    jpeg_huff_build(hd.dht[0][hd.comp[c].td], hd.dht[0][hd.comp[c].td] + 16, &tabs[2 * c]);
    jpeg_huff_build(hd.dht[1][hd.comp[c].ta], hd.dht[1][hd.comp[c].ta] + 16, &tabs[2 * c + 1]);
    for (int i = 0; i < hd.comp[c].h * hd.comp[c].v; ++i) { if (bpm >= 16) return 0; comp_of_block[bpm++] = (uint8_t)c; }
  }
  // sequential decoder
  JpegBits br;
  jpeg_bits_init(&br, f.data() + hd.scan_offset, f.data() + f.size());
  int pred[3] = {0, 0, 0};
  long long checksum = 0;
  for (long long m = 0; m < (long long)hd.mcus_x * hd.mcus_y; ++m) {
    if (hd.restart_interval && m && m % hd.restart_interval == 0) { jpeg_bits_restart(&br); pred[0] = pred[1] = pred[2] = 0; }
    for (int b = 0; b < bpm; ++b) {
      int16_t blk[64] = {0};
      const int c = comp_of_block[b];
      jpeg_decode_block(&br, tabs[2 * c], tabs[2 * c + 1], &pred[c], blk);
      uint8_t px[64];
      jpeg_idct_islow_8x8(blk, hd.quant[hd.comp[c].tq], px, 8);
      checksum += px[0] + px[63];
    }
  }
  // window decoder primitives over the unstuffed scan
  std::vector<uint8_t> clean;
  for (size_t i = hd.scan_offset; i < f.size(); ++i) {
    if (f[i] == 0xFF) { if (i + 1 < f.size() && f[i + 1] == 0x00) { clean.push_back(0xFF); ++i; continue; } break; }
    clean.push_back(f[i]);
  }
  const long long nbits = (long long)clean.size() * 8;
  for (long long start = 0; start < nbits; start += 1024) {
    JpegSyncState st = {start, 0, 0};
    const long long end = start + 1024 < nbits ? start + 1024 : nbits;
    while (st.pos < end) {
      int zz, value; bool done;
      jpeg_sync_step(clean.data(), nbits, tabs, comp_of_block, bpm, &st, &zz, &value, &done);
      if (zz < -1 || zz > 63 || st.b < 0 || st.b >= bpm || st.k < 0 || st.k > 63) { printf("state out of range\n"); abort(); }
      checksum += value;
    }
  }
  return (int)(checksum & 1);
}

int main(int argc, char** argv) {
  long long cases = 0, accepted = 0;
  int sink = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* fp = fopen(argv[a], "rb");
    if (!fp) { printf("cannot open %s\n", argv[a]); return 2; }
    std::vector<uint8_t> orig;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) orig.insert(orig.end(), tmp, tmp + n);
    fclose(fp);
    JpegHeader hd;
    std::string msg;
    if (!jpeg_parse_header(orig.data(), orig.size(), &hd, &msg)) { printf("%s: %s\n", argv[a], msg.c_str()); return 2; }
    // where the DHT segments are
    std::vector<size_t> dht;
    for (size_t p = 2; p + 4 <= hd.scan_offset;) {
      if (orig[p] != 0xFF) break;
      const size_t L = ((size_t)orig[p + 2] << 8) | orig[p + 3];
      if (orig[p + 1] == 0xC4) dht.push_back(p);
      p += 2 + L;
    }
    // 1. targeted: over-subscribed code lengths and absurd symbols in every DHT table
    for (size_t d : dht) {
      const size_t L = ((size_t)orig[d + 2] << 8) | orig[d + 3];
      for (int v = 0; v < 40; ++v) {
        std::vector<uint8_t> f = orig;
        const size_t tab = d + 5;                                   // first table's counts[16]
        if (v < 16) f[tab + (size_t)v] = (uint8_t)(3 + (v << 3));    // e.g. three 1-bit codes
        else if (v < 32) { for (size_t i = tab + 16; i < d + 2 + L; ++i) f[i] = (uint8_t)(rnd() | 0xF0); }   // symbols with huge categories
        else { for (int k = 0; k < 4; ++k) f[tab + rnd() % 16] = (uint8_t)rnd(); }
        sink ^= decode(f, &accepted); ++cases;
      }
    }
    // 2. random damage in the header and in the scan
    for (int it = 0; it < 400; ++it) {
      std::vector<uint8_t> f = orig;
      const int nmut = 1 + (int)(rnd() % 6);
      for (int k = 0; k < nmut; ++k) {
        const size_t pos = (it & 1) ? rnd() % hd.scan_offset : hd.scan_offset + rnd() % (f.size() - hd.scan_offset);
        f[pos] = (uint8_t)rnd();
      }
      if (it % 7 == 0) f.resize(hd.scan_offset + rnd() % (f.size() - hd.scan_offset));    // truncated scans
      sink ^= decode(f, &accepted); ++cases;
    }
  }
  printf("fuzzed %lld files, %lld accepted by the parser, no sanitizer report (%d)\n", cases, accepted, sink);
  return 0;
}
