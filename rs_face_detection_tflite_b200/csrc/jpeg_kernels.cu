// jpeg_kernels.cu -- see jpeg_device.h.  The three kernels of the device JPEG decoder (frame ingest, utils.rs:8-21).
#include "jpeg_device.h"
#include "glue_math.h"
#include "jpeg_color.cuh"

#include <atomic>
#include <climits>

namespace fdl {

void count_launch();

namespace {

template <typename K>
void opt_in_smem_once(std::atomic<unsigned long long>& done, K kernel, int bytes) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (done.load(std::memory_order_acquire) & bit) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done.fetch_or(bit, std::memory_order_release);
}

// ------------------------------------------------------------------------------------------------ entropy stage
__constant__ uint8_t kZigZag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kLookBits = 11;   // first-level Huffman look-up of the kernel (JpegHuff itself carries 9 bits)

struct EntropyShared {
  JpegHuff tabs[6];          // distinct Huffman tables of the image ("slots"; at most 2 per component)
  JpegImageDesc d;
  int blk_off[kJpegMaxBlocksPerMcu];  // coefficient offset of block b inside its MCU: (by * bcols + bx) * 64
  int mcu_step[3], row_jump[3], comp_rel[3];   // per component: next MCU / extra at the end of an MCU row / plane start relative to component 0
  uint32_t comp_pack;        // component of block b in bits 2b..2b+1
  uint32_t slot_pack;        // table slot of (component c, AC?) in bits 4(2c+a)..
  int nslots;
  int warp_part[32];
  int scan_total;
  int term;                  // first byte (relative to the aligned start) of the marker that ends the scan
  uint8_t zz[64];
};

// phase timestamps of CTA 0 (globaltimer ns): start, tables, compaction, round 0, hand-over rounds, scans, output; [7] = rounds
__device__ long long g_jpeg_phase[8];
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define JPEG_PHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_jpeg_phase[i] = gtime(); } while (0)

struct DecState { uint32_t pos; int b, k; };

__device__ __forceinline__ uint32_t be32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// Per-window records of the self-synchronising schedule (dynamic shared memory, nwin_cap entries each).
struct WindowArrays {
  unsigned long long* exit_;   // exit state: pos | (b << 8 | k) << 32
  uint32_t* epos;              // entry state
  uint32_t* ebk;
  int* nb;                     // blocks completed in the window
  int* dc;                     // [3][cap] sums of DC differences
  const uint16_t* look;        // [nslots][1 << kLookBits]: (length << 8) | symbol by the next 11 bits, 0 for longer codes
  int cap;
};

// The symbol loop -- jpeg_sync_step (jpeg_math.h) on a 64-bit register window of the clean scan (MSB first; the scan is followed
// by zero bytes).  The register window always ends on a 32-bit word boundary of the scan and holds >= 32 bits before every
// symbol, so a Huffman code (<= 16 bits) and its extra bits (<= 16) come from one peek, and a refill is one aligned word load per
// 32 consumed bits, requested one refill ahead (per-lane loads are what the L1 serialises, and their latency would sit in the
// symbol-to-symbol chain).  The body is straight-line (selects instead of branches: the lanes of a warp sit in different places of
// different blocks); codes longer than 11 bits are resolved by counting the left-aligned code limits they reach.
//   WRITE = false: windows i0 .. i1-1 are decoded back to back and recorded in `W` (entry / exit states, blocks, DC sums).
//   WRITE = true:  runs while pos < end and fewer than max_blocks blocks are complete; dc[] are the running predictors; coefficients
//                  go to their blocks, the first being block st.b of MCU `m`.
// Returns the number of blocks completed.
template <bool WRITE>
__device__ __forceinline__ int jpeg_run(const uint32_t* w, const EntropyShared& S, DecState& st, uint32_t end, int max_blocks, int dc[3],
                                        int16_t* coef, int m, const WindowArrays& W, int i0, int i1, uint32_t WB, uint32_t nbits) {
  const JpegImageDesc& d = S.d;
  uint32_t pos = st.pos;
  int b = st.b, k = st.k, nb = 0;
  int dc0 = dc[0], dc1 = dc[1], dc2 = dc[2];
  const uint32_t comp_pack = S.comp_pack, slot_pack = S.slot_pack;
  const int bpm = d.bpm;
  int c = (int)((comp_pack >> (2 * b)) & 3u);
  // WRITE: coefficient offsets (int16 elements, relative to component 0's plane) of the current MCU per component
  int mx = 0, base0 = 0, base1 = 0, base2 = 0, blk = 0;
  int16_t* cbase = nullptr;
  if (WRITE) {
    cbase = coef + d.coef_off[0];
    const int my = m / d.mcux;
    mx = m - my * d.mcux;
    base0 = S.comp_rel[0] + my * (S.mcu_step[0] * d.mcux + S.row_jump[0]) + mx * S.mcu_step[0];
    base1 = S.comp_rel[1] + my * (S.mcu_step[1] * d.mcux + S.row_jump[1]) + mx * S.mcu_step[1];
    base2 = S.comp_rel[2] + my * (S.mcu_step[2] * d.mcux + S.row_jump[2]) + mx * S.mcu_step[2];
    blk = (c == 0 ? base0 : (c == 1 ? base1 : base2)) + S.blk_off[b];
  }
  // window bookkeeping of the recording mode
  int wi_ = i0, nb_mark = 0, m0 = 0, m1 = 0, m2 = 0;
  uint32_t wend = end;
  if (!WRITE) {
    W.epos[i0] = pos; W.ebk[i0] = (uint32_t)(b << 8 | k);
    const unsigned long long e = (unsigned long long)(i0 + 1) * WB;
    wend = e < nbits ? (uint32_t)e : nbits;
  }
  uint32_t wi = pos >> 5;
  unsigned long long acc = (((unsigned long long)be32(w[wi]) << 32) | be32(w[wi + 1])) << (pos & 31);
  int n = 64 - (int)(pos & 31);
  wi += 2;
  uint32_t nxt = w[wi];                               // the next refill word, requested one refill ahead
  for (;;) {
    if (!WRITE) {
      if (pos >= wend) {                              // the window is complete: record it, open the next one with the same state
        W.exit_[wi_] = (unsigned long long)pos | ((unsigned long long)(b << 8 | k) << 32);
        W.nb[wi_] = nb - nb_mark; nb_mark = nb;
        W.dc[wi_] = dc0 - m0; W.dc[W.cap + wi_] = dc1 - m1; W.dc[2 * W.cap + wi_] = dc2 - m2;
        m0 = dc0; m1 = dc1; m2 = dc2;
        if (++wi_ >= i1) break;
        W.epos[wi_] = pos; W.ebk[wi_] = (uint32_t)(b << 8 | k);
        const unsigned long long e = (unsigned long long)(wi_ + 1) * WB;
        wend = e < nbits ? (uint32_t)e : nbits;
        continue;                                     // (a symbol may span more than one window)
      }
    } else {
      if (pos >= end || nb >= max_blocks) break;
    }
    const bool ac = k != 0;
    const int slot = (int)((slot_pack >> (4 * (2 * c + (ac ? 1 : 0)))) & 15u);
    const uint32_t v = (uint32_t)(acc >> 32);
    const int look = W.look[(slot << kLookBits) + (int)(v >> (32 - kLookBits))];
    int len = look >> 8, sym = look & 255;
    if (look == 0) {                                  // a code of 12..16 bits: its length = 12 + the number of limits it reaches
      const JpegHuff& t = S.tabs[slot];
      const uint32_t v16 = v >> 16;
      len = 12 + (v16 >= t.limit[12]) + (v16 >= t.limit[13]) + (v16 >= t.limit[14]) + (v16 >= t.limit[15]);
      sym = v16 >= t.limit[16] ? 0 : t.huffval[((int)(v >> (32 - len)) + t.valoffset[len]) & 255];   // no such code: libjpeg's zero symbol
    }
    const int s = ac ? (sym & 15) : min(sym, 16);     // extra bits
    const int kk = ac ? k + (sym >> 4) : 0;           // zig-zag position of the coefficient
    const bool has_val = s != 0 && kk <= 63;
    const int val = has_val ? jpeg_extend((int)((v << len) >> (32 - s)), s) : 0;
    const int used = len + (has_val ? s : 0);
    // the next position in the block: EOB -> 64, ZRL -> k + 16, a coefficient -> its position + 1 (a run past 63 ends the block)
    const int kn = ac ? (s ? kk + 1 : ((sym >> 4) == 15 ? k + 16 : 64)) : 1;
    if (!ac) { dc0 += c == 0 ? val : 0; dc1 += c == 1 ? val : 0; dc2 += c == 2 ? val : 0; }
    if (WRITE) {
      if (!ac) cbase[blk] = (int16_t)(c == 0 ? dc0 : (c == 1 ? dc1 : dc2));
      else if (has_val) cbase[blk + S.zz[kk]] = (int16_t)val;
    }
    pos += used; acc <<= used; n -= used;
    if (n < 32) {
      acc |= (unsigned long long)be32(nxt) << (32 - n); n += 32;
      ++wi;
      nxt = w[wi];
      if ((wi & 31u) == 8u) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + wi + 24));   // the next 128-byte line of the scan
    }
    // block complete?  (selects: some lane of the warp finishes a block in almost every iteration)
    const bool done = kn > 63;
    k = done ? 0 : kn;
    nb += done ? 1 : 0;
    const bool wrap = done && b + 1 == bpm;
    b = done ? (wrap ? 0 : b + 1) : b;
    c = (int)((comp_pack >> (2 * b)) & 3u);
    if (WRITE) {
      if (wrap) {
        base0 += S.mcu_step[0]; base1 += S.mcu_step[1]; base2 += S.mcu_step[2];
        if (++mx == d.mcux) { mx = 0; base0 += S.row_jump[0]; base1 += S.row_jump[1]; base2 += S.row_jump[2]; }
      }
      if (done) blk = (c == 0 ? base0 : (c == 1 ? base1 : base2)) + S.blk_off[b];
    }
  }
  st.pos = pos; st.b = b; st.k = k;
  dc[0] = dc0; dc[1] = dc1; dc[2] = dc2;
  return nb;
}

// exclusive block scan of one int per thread; *total = the sum.  Three barriers; `S.warp_part` is free again on return.
__device__ __forceinline__ int block_excl_scan(int v, EntropyShared& S, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) S.warp_part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int p = lane < nwarp ? S.warp_part[lane] : 0;
    int q = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, q, o); if (lane >= o) q += u; }
    S.warp_part[lane] = q - p;
    if (lane == 31) S.scan_total = q;
  }
  __syncthreads();
  const int excl = incl - v + S.warp_part[warp];
  *total = S.scan_total;
  __syncthreads();
  return excl;
}

// ---- the scan without stuffing and restart markers (a stream compaction over all images of the batch at once):
// FF00 -> FF, FFDn dropped and remembered as the start of a restart interval, FFFF fill dropped, any other marker ends the scan.
// Tiles of 4 KB (256 threads x 16 bytes); pass 1 counts per tile, pass 2 sums the counts of the tiles before it, scans its own
// bytes and writes.  tile_info[img][tile] = {bytes kept, RSTn markers, first scan-ending marker (byte offset from the image's
// aligned start) or INT_MAX}; scan_len[img] = {bytes of the clean scan, RSTn markers in it}.
constexpr int kScanThreads = 256, kScanTile = kScanThreads * 16;

struct ScanMasks { uint32_t q[4]; uint32_t keep, rst; int term; };

__device__ __forceinline__ ScanMasks scan_masks(const uint8_t* __restrict__ bytes, long long lo_abs, long long hi_abs, long long a0, long long abs) {
  ScanMasks r;
  r.q[0] = r.q[1] = r.q[2] = r.q[3] = 0; r.keep = r.rst = 0; r.term = INT_MAX;
  int prev = 0, next = 0xD9;
  if (abs < hi_abs) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(bytes + abs));
    r.q[0] = u.x; r.q[1] = u.y; r.q[2] = u.z; r.q[3] = u.w;
    if (abs - 1 >= lo_abs) prev = __ldg(bytes + abs - 1);
    if (abs + 16 < hi_abs) next = __ldg(bytes + abs + 16);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = (r.q[j >> 2] >> (8 * (j & 3))) & 255;
    const int nx = j < 15 ? (int)((r.q[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 255) : next;
    const long long pa = abs + j;
    if (pa >= lo_abs && pa < hi_abs) {
      const int nxe = pa + 1 < hi_abs ? nx : 0xD9;
      if (c == 0xFF) {
        if (nxe == 0) r.keep |= 1u << j;
        else if (nxe >= 0xD0 && nxe <= 0xD7) r.rst |= 1u << j;
        else if (nxe != 0xFF && r.term == INT_MAX) r.term = (int)(pa - a0);
      } else if (!(prev == 0xFF && (c == 0 || (c >= 0xD0 && c <= 0xD7)))) {
        r.keep |= 1u << j;
      }
    }
    prev = c;
  }
  return r;
}
// bytes at or behind the scan-ending marker do not count
__device__ __forceinline__ void scan_cut(ScanMasks& m, long long a0, int term, long long abs) {
  if (term == INT_MAX) return;
  const long long first_dead = a0 + term - abs;
  const uint32_t alive = first_dead >= 16 ? 0xFFFFu : (first_dead <= 0 ? 0u : ((1u << (int)first_dead) - 1u));
  m.keep &= alive; m.rst &= alive;
}

__global__ void __launch_bounds__(kScanThreads) jpeg_scan_count_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ bytes,
                                                                        int* __restrict__ tile_info, int max_tiles) {
  __shared__ int s_term, s_cnt[kScanThreads / 32];
  const JpegImageDesc& d = descs[blockIdx.y];
  const long long lo_abs = d.raw_off, hi_abs = d.raw_off + d.raw_len, a0 = lo_abs & ~15LL;
  const long long tile0 = a0 + (long long)blockIdx.x * kScanTile;
  if (tile0 >= hi_abs) return;
  const int tid = threadIdx.x;
  if (tid == 0) s_term = INT_MAX;
  __syncthreads();
  const long long abs = tile0 + (long long)tid * 16;
  ScanMasks m = scan_masks(bytes, lo_abs, hi_abs, a0, abs);
  if (m.term != INT_MAX) atomicMin(&s_term, m.term);
  __syncthreads();
  const int term = s_term;
  scan_cut(m, a0, term, abs);
  int v = __popc(m.keep) | (__popc(m.rst) << 16);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) t += s_cnt[i];
    int* o = tile_info + ((long long)blockIdx.y * max_tiles + blockIdx.x) * 3;
    o[0] = t & 0xFFFF; o[1] = t >> 16; o[2] = term;
  }
}

__global__ void __launch_bounds__(kScanThreads) jpeg_scan_compact_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ bytes,
                                                                          const int* __restrict__ tile_info, int max_tiles, uint8_t* __restrict__ clean_all,
                                                                          int* __restrict__ iv_all, int* __restrict__ scan_len) {
  __shared__ int s_red[3][kScanThreads / 32], s_base[3], s_warp[kScanThreads / 32];
  const JpegImageDesc& d = descs[blockIdx.y];
  const long long lo_abs = d.raw_off, hi_abs = d.raw_off + d.raw_len, a0 = lo_abs & ~15LL;
  const long long tile0 = a0 + (long long)blockIdx.x * kScanTile;
  if (tile0 >= hi_abs) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (int)((hi_abs - a0 + kScanTile - 1) / kScanTile);
  const int* info = tile_info + (long long)blockIdx.y * max_tiles * 3;
  // the tiles before this one: kept bytes, markers, and whether the scan already ended there
  int pb = 0, pr = 0, dead = 0;
  for (int t = tid; t < (int)blockIdx.x; t += kScanThreads) { pb += info[3 * t]; pr += info[3 * t + 1]; dead |= info[3 * t + 2] != INT_MAX; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { pb += __shfl_xor_sync(0xffffffffu, pb, o); pr += __shfl_xor_sync(0xffffffffu, pr, o); dead |= __shfl_xor_sync(0xffffffffu, dead, o); }
  if (lane == 0) { s_red[0][warp] = pb; s_red[1][warp] = pr; s_red[2][warp] = dead; }
  __syncthreads();
  if (tid == 0) {
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; c |= s_red[2][i]; }
    s_base[0] = a; s_base[1] = b; s_base[2] = c;
  }
  __syncthreads();
  if (s_base[2]) return;                             // an earlier tile holds the marker that ends the scan
  const int term = info[3 * blockIdx.x + 2];
  const long long abs = tile0 + (long long)tid * 16;
  ScanMasks m = scan_masks(bytes, lo_abs, hi_abs, a0, abs);
  scan_cut(m, a0, term, abs);
  const int v = __popc(m.keep) | (__popc(m.rst) << 16);
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int i = 0; i < warp; ++i) wbase += s_warp[i];
  const int excl = incl - v + wbase;
  uint8_t* clean = clean_all + d.clean_off;
  int* ivs = iv_all + d.iv_off;
  int o = s_base[0] + (excl & 0xFFFF), r = s_base[1] + (excl >> 16);
  if (m.keep | m.rst) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (m.keep & (1u << j)) clean[o++] = (uint8_t)((m.q[j >> 2] >> (8 * (j & 3))) & 255);
      else if (m.rst & (1u << j)) { if (++r < d.n_intervals) ivs[r] = o; }
    }
  }
  // the last live tile closes the scan: its length, the marker count, 64 zero bytes behind it, interval 0
  if (term != INT_MAX || (int)blockIdx.x == ntiles - 1) {
    int tb = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) tb += s_warp[i];
    const int len = s_base[0] + (tb & 0xFFFF);
    if (tid < 64) clean[len + tid] = 0;
    if (tid == 0) {
      scan_len[2 * blockIdx.y] = len; scan_len[2 * blockIdx.y + 1] = s_base[1] + (tb >> 16);
      if (d.n_intervals > 0) ivs[0] = 0;
    }
  }
}

__global__ void __launch_bounds__(kJpegEntropyThreads, 2)
jpeg_entropy_kernel(const JpegImageDesc* __restrict__ descs, const JpegHuff* __restrict__ tabs, const int* __restrict__ scan_len,
                    const uint8_t* clean_all, int16_t* coef, const int* iv_all, int* status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EntropyShared& S = *reinterpret_cast<EntropyShared*>(smem_raw);
  const int tid = threadIdx.x, T = blockDim.x, img = blockIdx.x;
  JPEG_PHASE(0);
  // ---- descriptor, tables
  {
    const int* src = reinterpret_cast<const int*>(descs + img);
    int* dst = reinterpret_cast<int*>(&S.d);
    for (int i = tid; i < (int)(sizeof(JpegImageDesc) / 4); i += T) dst[i] = src[i];
  }
  __syncthreads();
  const JpegImageDesc& d = S.d;
  if (tid == 0) {
    uint32_t pack = 0;
    for (int b = 0; b < d.bpm; ++b) { pack |= (uint32_t)d.blk_comp[b] << (2 * b); S.blk_off[b] = (d.blk_by[b] * d.bcols[d.blk_comp[b]] + d.blk_bx[b]) * 64; }
    S.comp_pack = pack;
    for (int c = 0; c < 3; ++c) {
      const bool on = c < d.ncomp;
      S.mcu_step[c] = on ? d.hs[c] * 64 : 0;
      S.row_jump[c] = on ? (d.vs[c] - 1) * d.bcols[c] * 64 : 0;
      S.comp_rel[c] = on ? (int)(d.coef_off[c] - d.coef_off[0]) : 0;
    }
    // distinct tables -> slots (Cb and Cr normally share theirs)
    int ids[6], ns = 0;
    uint32_t sp = 0;
    for (int c = 0; c < d.ncomp; ++c)
      for (int a = 0; a < 2; ++a) {
        const int id = a ? d.tab_ac[c] : d.tab_dc[c];
        int sl = 0;
        while (sl < ns && ids[sl] != id) ++sl;
        if (sl == ns) ids[ns++] = id;
        sp |= (uint32_t)sl << (4 * (2 * c + a));
      }
    S.slot_pack = sp; S.nslots = ns;
    for (int sl = 0; sl < ns; ++sl) S.warp_part[sl] = ids[sl];
  }
  if (tid < 64) S.zz[tid] = kZigZag[tid];
  __syncthreads();
  const int nslots = S.nslots;
  for (int sl = 0; sl < nslots; ++sl) {
    const int* src = reinterpret_cast<const int*>(tabs + S.warp_part[sl]);
    int* dst = reinterpret_cast<int*>(&S.tabs[sl]);
    for (int i = tid; i < (int)(sizeof(JpegHuff) / 4); i += T) dst[i] = src[i];
  }
  __syncthreads();
  unsigned char* dyn = smem_raw + ((sizeof(EntropyShared) + 15) & ~size_t(15));
  uint16_t* s_look = reinterpret_cast<uint16_t*>(dyn);
  // 11-bit first-level look-ups from the 9-bit ones + the code limits
  for (int e = tid; e < nslots * (1 << kLookBits); e += T) {
    const int sl = e >> kLookBits, idx = e & ((1 << kLookBits) - 1);
    const JpegHuff& t = S.tabs[sl];
    int look = t.look[idx >> (kLookBits - 9)];
    if (look == 0) {
      const uint32_t v16 = (uint32_t)idx << (16 - kLookBits);
      const int len = 10 + (v16 >= t.limit[10]);
      if (v16 < t.limit[len]) look = (len << 8) | t.huffval[((idx >> (kLookBits - len)) + t.valoffset[len]) & 255];
    }
    s_look[e] = (uint16_t)look;
  }
  __syncthreads();

  const uint8_t* clean = clean_all + d.clean_off;
  const int* ivs = iv_all + d.iv_off;
  const int run_bytes = scan_len[2 * img], run_rst = scan_len[2 * img + 1];
  JPEG_PHASE(1);
  JPEG_PHASE(2);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(clean);
  const uint32_t nbits = (uint32_t)run_bytes * 8u;
  const int total_blocks = d.mcux * d.mcuy * d.bpm;
  WindowArrays W;
  W.look = s_look;
  W.cap = d.nwin_cap;
  W.exit_ = reinterpret_cast<unsigned long long*>(dyn + (((size_t)nslots << (kLookBits + 1)) + 15 & ~size_t(15)));
  W.epos = reinterpret_cast<uint32_t*>(W.exit_ + W.cap);
  W.ebk = W.epos + W.cap;
  W.nb = reinterpret_cast<int*>(W.ebk + W.cap);
  W.dc = W.nb + W.cap;

  // ---- 2a. restart intervals: byte-aligned, predictors reset, position known from the interval index: one thread each
  if (d.restart_interval > 0) {
    const int mcus = d.mcux * d.mcuy;
    const bool ok = run_rst + 1 >= d.n_intervals;
    if (ok) {
      for (int iv = tid; iv < d.n_intervals; iv += T) {
        DecState st = {(uint32_t)ivs[iv] * 8u, 0, 0};
        const int m0 = iv * d.restart_interval, m1 = min(m0 + d.restart_interval, mcus);
        int pred[3] = {0, 0, 0};
        jpeg_run<true>(w, S, st, nbits, (m1 - m0) * d.bpm, pred, coef, m0, W, 0, 0, 0u, nbits);
      }
    }
    if (tid == 0) status[img] = ok ? JPEG_OK : JPEG_ERR_RESTARTS;
    return;
  }

  // ---- 2b. no restart markers: self-synchronising windows (Weissenberger & Schmidt, ICPP 2018; jpeg_math.h)
  const int WB = d.window_bits;
  const int nwin = min((int)((nbits + (uint32_t)WB - 1u) / (uint32_t)WB), d.nwin_cap);
  const int per = (nwin + T - 1) / T;
  const int wlo = min(tid * per, nwin), whi = min(wlo + per, nwin);
  int none[3] = {0, 0, 0};
  // round 0: the first window of every thread starts from a guess (block 0, DC symbol next, at the window's first bit -- true for
  // window 0); the thread's other windows follow in the same symbol loop
  if (wlo < whi) {
    DecState st = {(uint32_t)((unsigned long long)wlo * (unsigned)WB), 0, 0};
    jpeg_run<false>(w, S, st, 0u, INT_MAX, none, nullptr, 0, W, wlo, whi, (uint32_t)WB, nbits);
  }
  // hand-over rounds: a window whose predecessor's exit state differs from the entry state it was decoded from is decoded again
  // (and so on down the thread's windows until an exit state comes out as before); a fixed point is the sequential decode.  Exit
  // states are snapshotted between two barriers, so a round reads only the previous round's values.
  __syncthreads();
  JPEG_PHASE(3);
  int rounds = 0;
  for (;;) {
    ++rounds;
    unsigned long long ex = 0;
    const bool have = wlo > 0 && wlo < whi;
    if (have) ex = W.exit_[wlo - 1];
    __syncthreads();
    int changed = 0;
    if (have) {
      for (int i = wlo; i < whi; ++i) {
        if (i > wlo) ex = W.exit_[i - 1];
        const uint32_t pos = (uint32_t)ex, bk = (uint32_t)(ex >> 32) & 0xFFFFu;
        if (pos == W.epos[i] && bk == W.ebk[i]) break;
        DecState st = {pos, (int)(bk >> 8), (int)(bk & 255)};
        none[0] = none[1] = none[2] = 0;
        jpeg_run<false>(w, S, st, 0u, INT_MAX, none, nullptr, 0, W, i, i + 1, (uint32_t)WB, nbits);
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  JPEG_PHASE(4);
  if (blockIdx.x == 0 && tid == 0) g_jpeg_phase[7] = rounds;
  // block and DC prefix sums over the windows place every thread's output
  int my_nb = 0, my_dc[3] = {0, 0, 0};
  for (int i = wlo; i < whi; ++i) {
    my_nb += W.nb[i];
    my_dc[0] += W.dc[i]; my_dc[1] += W.dc[W.cap + i]; my_dc[2] += W.dc[2 * W.cap + i];
  }
  int tot_nb, tot;
  const int g = block_excl_scan(my_nb, S, &tot_nb);
  int pred[3];
  pred[0] = block_excl_scan(my_dc[0], S, &tot);
  pred[1] = block_excl_scan(my_dc[1], S, &tot);
  pred[2] = block_excl_scan(my_dc[2], S, &tot);
  JPEG_PHASE(5);
  // output pass: the thread's windows once more, from the (now true) entry state of its first one, writing coefficients and
  // absolute DC values
  if (wlo < whi && g < total_blocks) {
    DecState st = {W.epos[wlo], (int)(W.ebk[wlo] >> 8), (int)(W.ebk[wlo] & 255)};
    const unsigned long long e = (unsigned long long)whi * (unsigned)WB;
    jpeg_run<true>(w, S, st, e < nbits ? (uint32_t)e : nbits, total_blocks - g, pred, coef, g / d.bpm, W, 0, 0, 0u, nbits);
  }
  __syncthreads();
  JPEG_PHASE(6);
  if (tid == 0) status[img] = tot_nb >= total_blocks ? JPEG_OK : JPEG_ERR_BLOCKS;
}

std::atomic<unsigned long long> g_entropy_optin{0};

// ------------------------------------------------------------------------------------------------ dequantise + IDCT
// range_limit[x & 0x3FF] of jdmaster.c (the table is centred on +128) == clamp(sign-extended low 10 bits of x + 128, 0, 255)
__device__ __forceinline__ uint32_t range_limit_dev(int x) {
  const int s = (x << 22) >> 22;
  return (uint32_t)min(max(s + 128, 0), 255);
}

// One warp = four horizontally adjacent blocks of one component; lane (j, t) loads row t of block j (16 bytes), runs column t,
// then row t, and stores 8 samples: the four blocks' rows make whole 32-byte sectors of the plane.  Two warp-uniform short cuts
// with the same results as the full transform (jidctint.c has the column one): all four blocks DC-only -> the block is the
// constant (dc * q + 4) >> 3; all four blocks with nothing below their first row -> the column pass is a shift.
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegImageDesc* __restrict__ descs, const int16_t* __restrict__ coef,
                                                        uint8_t* __restrict__ planes) {
  __shared__ int ws[8][4][8][9];
  __shared__ uint16_t s_quant[3][64];
  __shared__ long long s_coef_off[3], s_plane_off[3];
  __shared__ int s_bcols[3], s_qpr[3], s_qcount[3], s_ncomp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = lane >> 3, t = lane & 7;
  {
    const JpegImageDesc& d = descs[blockIdx.y];
    if (tid < 192) s_quant[tid >> 6][tid & 63] = d.quant[tid >> 6][tid & 63];
    if (tid < 3) {
      s_coef_off[tid] = d.coef_off[tid]; s_plane_off[tid] = d.plane_off[tid]; s_bcols[tid] = d.bcols[tid];
      s_qpr[tid] = (d.bcols[tid] + 3) >> 2; s_qcount[tid] = tid < d.ncomp ? d.brows[tid] * ((d.bcols[tid] + 3) >> 2) : 0;
    }
    if (tid == 0) s_ncomp = d.ncomp;
  }
  __syncthreads();
  int q = blockIdx.x * 8 + warp, c = 0;
  for (; c < s_ncomp; ++c) {
    if (q < s_qcount[c]) break;
    q -= s_qcount[c];
  }
  if (c >= s_ncomp) return;
  const int bcols = s_bcols[c], qpr = s_qpr[c], row = q / qpr, bx = (q - row * qpr) * 4 + j;
  const bool valid = bx < bcols;
  int (*m)[9] = ws[warp][j];
  int v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0;
  if (valid) {
    const int4 u = __ldg(reinterpret_cast<const int4*>(coef + s_coef_off[c] + ((long long)row * bcols + bx) * 64 + t * 8));
    const int p[4] = {u.x, u.y, u.z, u.w};
    const uint4 qa = *reinterpret_cast<const uint4*>(&s_quant[c][t * 8]);
    const uint32_t qq[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = (int)(int16_t)(p[i] & 0xFFFF) * (int)(qq[i] & 0xFFFFu);
      v[2 * i + 1] = (p[i] >> 16) * (int)(qq[i] >> 16);
    }
  }
  uint8_t* dst = planes + s_plane_off[c] + (long long)(row * 8 + t) * (bcols * 8) + bx * 8;
  const int ac_in_row = v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7];
  const unsigned below = __ballot_sync(0xffffffffu, t > 0 && (ac_in_row | v[0]) != 0);
  int x[8];
  if (below == 0) {
    const unsigned ac0 = __ballot_sync(0xffffffffu, t == 0 && ac_in_row != 0);
    if (ac0 == 0) {                                  // DC only: ((dc*q << 2) << 13 + 2^17) >> 18
      const int dcq = __shfl_sync(0xffffffffu, v[0], lane & 24);
      const uint32_t b = range_limit_dev((dcq + 4) >> 3) * 0x01010101u;
      if (valid) *reinterpret_cast<uint2*>(dst) = make_uint2(b, b);
      return;
    }
    // only the first row is populated: every column's pass yields row0[c] << PASS1_BITS in all eight rows
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (int)((unsigned)__shfl_sync(0xffffffffu, v[i], lane & 24) << 2);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[t][i] = v[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = m[i][t];
    jpeg_idct_1d(x, 1, 13, 13 - 2);                                // column t
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i][t] = x[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = m[t][i];
  }
  jpeg_idct_1d(x, 1, 13, 13 + 2 + 3);                              // row t
  if (valid) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lo |= range_limit_dev(x[i]) << (8 * i);
      hi |= range_limit_dev(x[4 + i]) << (8 * i);
    }
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}

// One THREAD per block: the 64 dequantised coefficients live in registers, eight column passes and eight row passes run on
// compile-time indices (no shared-memory transposes, no shuffles), and a column / row whose AC terms are all zero takes jidctint.c's
// short cut (identical results: DESCALE of a multiple of 2^13).  Consecutive threads take consecutive blocks of a block row, so the
// eight 8-byte row stores of a warp are 256 contiguous bytes each.  20 executed instructions per coefficient against the 36 of the
// eight-lanes-per-block kernel above (which stays for A/B: FDL_JPEG_IDCT=0).
// The kernel also hands the coefficient buffer back CLEAN: every 16-byte row that held a non-zero coefficient is zeroed after it
// has been read (the entropy stage writes non-zero coefficients only, into a zeroed buffer), which replaces a 1.6 GB memset per 256
// 1080p frames with a few predicated stores.
__global__ void __launch_bounds__(128) jpeg_idct_block_kernel(const JpegImageDesc* __restrict__ descs, int16_t* __restrict__ coef,
                                                              uint8_t* __restrict__ planes) {
  __shared__ uint16_t s_quant[3][64];
  __shared__ long long s_coef_off[3], s_plane_off[3];
  __shared__ int s_bcols[3], s_count[3], s_ncomp;
  const int tid = threadIdx.x;
  {
    const JpegImageDesc& d = descs[blockIdx.y];
    for (int i = tid; i < 192; i += 128) s_quant[i >> 6][i & 63] = d.quant[i >> 6][i & 63];
    if (tid < 3) {
      s_coef_off[tid] = d.coef_off[tid]; s_plane_off[tid] = d.plane_off[tid]; s_bcols[tid] = d.bcols[tid];
      s_count[tid] = tid < d.ncomp ? d.brows[tid] * d.bcols[tid] : 0;
    }
    if (tid == 0) s_ncomp = d.ncomp;
  }
  __syncthreads();
  int g = blockIdx.x * 128 + tid, c = 0;
  for (; c < s_ncomp; ++c) {
    if (g < s_count[c]) break;
    g -= s_count[c];
  }
  if (c >= s_ncomp) return;
  const int bcols = s_bcols[c], row = g / bcols, bx = g - row * bcols;
  int4* src = reinterpret_cast<int4*>(coef + s_coef_off[c] + (long long)g * 64);
  int ws[64];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int4 u = src[r];
    if ((u.x | u.y | u.z | u.w) != 0) src[r] = make_int4(0, 0, 0, 0);
    const int p[4] = {u.x, u.y, u.z, u.w};
    const uint4 qa = *reinterpret_cast<const uint4*>(&s_quant[c][r * 8]);
    const uint32_t qq[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ws[8 * r + 2 * i] = (int)(int16_t)(p[i] & 0xFFFF) * (int)(qq[i] & 0xFFFFu);
      ws[8 * r + 2 * i + 1] = (p[i] >> 16) * (int)(qq[i] >> 16);
    }
  }
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    const int ac = ws[8 + col] | ws[16 + col] | ws[24 + col] | ws[32 + col] | ws[40 + col] | ws[48 + col] | ws[56 + col];
    if (ac == 0) {
      const int dc = (int)((unsigned)ws[col] << 2);       // PASS1_BITS
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[8 * r + col] = dc;
    } else {
      jpeg_idct_1d(ws + col, 8, 13, 13 - 2);
    }
  }
  uint8_t* dst = planes + s_plane_off[c] + (long long)(row * 8) * (bcols * 8) + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int* x = ws + 8 * r;
    uint32_t lo, hi;
    if ((x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0) {
      const uint32_t b = range_limit_dev((x[0] + 16) >> 5) * 0x01010101u;   // DESCALE(x0 << 13, 13 + PASS1_BITS + 3)
      lo = hi = b;
    } else {
      jpeg_idct_1d(x, 1, 13, 13 + 2 + 3);
      lo = range_limit_dev(x[0]) | range_limit_dev(x[1]) << 8 | range_limit_dev(x[2]) << 16 | range_limit_dev(x[3]) << 24;
      hi = range_limit_dev(x[4]) | range_limit_dev(x[5]) << 8 | range_limit_dev(x[6]) << 16 | range_limit_dev(x[7]) << 24;
    }
    *reinterpret_cast<uint2*>(dst + (long long)r * (bcols * 8)) = make_uint2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------------ upsampling + colour conversion
// Generic path (one pixel group at a time through jdsample.c's gather form): any mix of 1x1 / 2x1 / 2x2 components, components at
// most 2 samples wide (plain replication), unaligned output.  Images the fast kernel below takes are skipped.
__device__ __forceinline__ void chroma4(const uint8_t* __restrict__ plane, int stride, int cw, int ch, int eh, int ev, int x0, int y, int out[4]) {
  if (eh == 1) {
    const uint8_t* r = plane + (long long)y * stride + x0;
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = r[i];      // the plane is MCU-padded: x0 + 3 stays inside the row
    return;
  }
  const int cx = x0 >> 1;
  if (ev == 1) {                                     // h2v1
    const uint8_t* r = plane + (long long)y * stride;
    const int s0 = r[cx], s1 = r[cx + 1];            // cx + 1 < padded width (a multiple of 8)
    if (cw <= 2) { out[0] = out[1] = s0; out[2] = out[3] = s1; return; }
    const int sm = cx > 0 ? r[cx - 1] : 0, s2 = cx + 2 < cw ? r[cx + 2] : 0;
    out[0] = cx == 0 ? s0 : (3 * s0 + sm + 1) >> 2;
    out[1] = cx == cw - 1 ? s0 : (3 * s0 + s1 + 2) >> 2;
    out[2] = (3 * s1 + s0 + 1) >> 2;
    out[3] = cx + 1 >= cw - 1 ? s1 : (3 * s1 + s2 + 2) >> 2;
    return;
  }
  const int cy = y >> 1;                             // h2v2
  if (cw <= 2) {
    const uint8_t* r = plane + (long long)cy * stride;
    out[0] = out[1] = r[cx]; out[2] = out[3] = r[cx + 1];
    return;
  }
  int fy = (y & 1) ? cy + 1 : cy - 1;
  fy = fy < 0 ? 0 : (fy > ch - 1 ? ch - 1 : fy);
  const uint8_t* nr = plane + (long long)cy * stride;
  const uint8_t* fr = plane + (long long)fy * stride;
  const int s0 = 3 * nr[cx] + fr[cx], s1 = 3 * nr[cx + 1] + fr[cx + 1];
  const int sm = cx > 0 ? 3 * nr[cx - 1] + fr[cx - 1] : 0, s2 = cx + 2 < cw ? 3 * nr[cx + 2] + fr[cx + 2] : 0;
  out[0] = cx == 0 ? (4 * s0 + 8) >> 4 : (3 * s0 + sm + 8) >> 4;
  out[1] = cx == cw - 1 ? (4 * s0 + 7) >> 4 : (3 * s0 + s1 + 7) >> 4;
  out[2] = (3 * s1 + s0 + 8) >> 4;
  out[3] = cx + 1 >= cw - 1 ? (4 * s1 + 7) >> 4 : (3 * s1 + s2 + 7) >> 4;
}

__global__ void __launch_bounds__(256) jpeg_color_generic_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                                                 uint8_t* __restrict__ out) {
  const JpegImageDesc& d = descs[blockIdx.z];
  if (d.color_fast) return;
  const int x0 = 4 * (blockIdx.x * 256 + threadIdx.x), y = blockIdx.y;
  if (x0 >= d.width || y >= d.height) return;
  const uint8_t* py = planes + d.plane_off[0] + (long long)y * (d.bcols[0] * 8) + x0;
  int Y[4], cb[4] = {128, 128, 128, 128}, cr[4] = {128, 128, 128, 128};
#pragma unroll
  for (int i = 0; i < 4; ++i) Y[i] = py[i];
  if (d.ncomp == 3) {
    chroma4(planes + d.plane_off[1], d.bcols[1] * 8, d.cw[1], d.ch[1], d.hmax / d.hs[1], d.vmax / d.vs[1], x0, y, cb);
    chroma4(planes + d.plane_off[2], d.bcols[2] * 8, d.cw[2], d.ch[2], d.hmax / d.hs[2], d.vmax / d.vs[2], x0, y, cr);
  }
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (d.ncomp == 1) px[3 * i] = px[3 * i + 1] = px[3 * i + 2] = (uint8_t)Y[i];     // IMREAD_COLOR of a greyscale file
    else jpeg_ycc_to_rgb(Y[i], cb[i], cr[i], px + 3 * i);
  }
  uint8_t* o = out + d.out_off + (long long)y * d.out_stride + 3LL * x0;
  const int n = min(4, d.width - x0) * 3;
  for (int i = 0; i < n; ++i) o[i] = px[i];
}

// Fast path: the usual files (colour with both chroma components at 2x2 or at 2x1, more than 2 chroma samples wide) into 16-byte
// aligned rows.  A warp walks 128 pixels x 2 rows at a time down 2 * kColorPairs rows (one chroma row + its two neighbours per
// pair, carried over to the next pair), a lane 4 pixels x 2 rows: aligned word loads, the near-row products shared by both rows,
// RGB staged through shared memory so that each lane stores 16 aligned bytes.
constexpr int kColorPairs = 4;     // row pairs a warp of jpeg_color_kernel walks down (a CTA: 256 pixels x 8 * kColorPairs rows)

__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                                         uint8_t* __restrict__ out) {
  __shared__ ColorShared P;
  __shared__ __align__(16) uint32_t s_rgb[8][2][96];
  {
    const JpegImageDesc& d = descs[blockIdx.z];
    if (!d.color_fast) return;
    if (threadIdx.x == 0) {
      P.py = planes + d.plane_off[0]; P.pcb = planes + d.plane_off[1]; P.pcr = planes + d.plane_off[2];
      P.out = out + d.out_off;
      P.sy = d.bcols[0] * 8; P.sc = d.bcols[1] * 8; P.width = d.width; P.height = d.height; P.out_stride = d.out_stride;
      P.cw = d.cw[1]; P.ch = d.ch[1]; P.v2 = d.vmax / d.vs[1] == 2;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xw = blockIdx.x * 256 + (warp & 1) * 128, x0 = xw + 4 * lane;     // first pixel of the warp / of the lane
  const int y0 = blockIdx.y * (8 * kColorPairs) + (warp >> 1) * (2 * kColorPairs);     // the warp's first row (even)
  if (xw >= P.width || y0 >= P.height) return;
  const bool active = x0 < P.width;
  const bool v2 = P.v2 != 0;
  const int cx = x0 >> 1, cw = P.cw, sy = P.sy, sc = P.sc, chm1 = P.ch - 1;
  // jdsample.c's fancy upsampling as dot products over the row words (bytes: samples c-1, c, c+1, c+2).  Far-row weights of the
  // four outputs; the near row takes three times them (h2v2) or is the only row (h2v1).  The image edges replace the missing
  // neighbour by the sample itself: weight 4 on it.  The rounding constant carries -128 (scaled), so the shift yields Cb - 128.
  const uint32_t k0 = cx == 0 ? 0x00000400u : 0x00000301u, k1 = cx == cw - 1 ? 0x00000400u : 0x00010300u, k2 = 0x00030100u,
                 k3 = cx + 1 >= cw - 1 ? 0x00040000u : 0x01030000u;
  const uint8_t* py = P.py + (long long)y0 * sy + x0;
  const int cy0 = v2 ? y0 >> 1 : y0;
  // h2v2: the chroma rows above / at / below the row pair, carried from pair to pair (one new row per pair and component)
  uint32_t w_up[2] = {0, 0}, w_at[2] = {0, 0}, w_dn[2] = {0, 0};
  if (active && v2) {
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
      const uint8_t* pl = comp ? P.pcr : P.pcb;
      w_up[comp] = row4(pl + (long long)max(cy0 - 1, 0) * sc, cx);
      w_at[comp] = row4(pl + (long long)cy0 * sc, cx);
      w_dn[comp] = row4(pl + (long long)min(cy0 + 1, chm1) * sc, cx);
    }
  }
  const int row_bytes = 3 * min(128, P.width - xw);
  uint8_t* orow = P.out + (long long)y0 * P.out_stride + 3LL * xw + 16 * lane;
  const int npairs = min(kColorPairs, (P.height - y0 + 1) >> 1);
  for (int j = 0; j < npairs; ++j) {
    uint32_t rgb[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (active) {
      const uint32_t ya = __ldg(reinterpret_cast<const uint32_t*>(py)), yb = __ldg(reinterpret_cast<const uint32_t*>(py + sy));
      int xbv[2][4], xrv[2][4];
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        const uint8_t* pl = comp ? P.pcr : P.pcb;
        int (*o)[4] = comp ? xrv : xbv;
        if (v2) {
          const uint32_t wn = w_at[comp];
          const uint32_t b8 = (uint32_t)(8 - 128 * 16), b7 = (uint32_t)(7 - 128 * 16);
          const uint32_t n0 = __dp4a(wn, 3u * k0, b8), n1 = __dp4a(wn, 3u * k1, b7), n2 = __dp4a(wn, 3u * k2, b8), n3 = __dp4a(wn, 3u * k3, b7);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t wf = r ? w_dn[comp] : w_up[comp];
            o[r][0] = (int)__dp4a(wf, k0, n0) >> 4;
            o[r][1] = (int)__dp4a(wf, k1, n1) >> 4;
            o[r][2] = (int)__dp4a(wf, k2, n2) >> 4;
            o[r][3] = (int)__dp4a(wf, k3, n3) >> 4;
          }
          w_up[comp] = wn; w_at[comp] = w_dn[comp];
          if (j + 1 < npairs) w_dn[comp] = row4(pl + (long long)min(cy0 + j + 2, chm1) * sc, cx);
        } else {
          const uint32_t b1 = (uint32_t)(1 - 128 * 4), b2 = (uint32_t)(2 - 128 * 4);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t wn = row4(pl + (long long)(cy0 + 2 * j + r) * sc, cx);
            o[r][0] = (int)__dp4a(wn, k0, b1) >> 2;
            o[r][1] = (int)__dp4a(wn, k1, b2) >> 2;
            o[r][2] = (int)__dp4a(wn, k2, b1) >> 2;
            o[r][3] = (int)__dp4a(wn, k3, b2) >> 2;
          }
        }
      }
      ycc_px4(ya, xbv[0], xrv[0], rgb[0]);
      ycc_px4(yb, xbv[1], xrv[1], rgb[1]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      s_rgb[warp][r][3 * lane] = rgb[r][0]; s_rgb[warp][r][3 * lane + 1] = rgb[r][1]; s_rgb[warp][r][3 * lane + 2] = rgb[r][2];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (y0 + 2 * j + r >= P.height || lane >= 24) continue;
      uint8_t* o = orow + (long long)r * P.out_stride;
      if (16 * lane + 16 <= row_bytes) *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(&s_rgb[warp][r][4 * lane]);
      else {
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(&s_rgb[warp][r][4 * lane]);
        for (int i = 0; 16 * lane + i < row_bytes; ++i) o[i] = sb[i];
      }
    }
    __syncwarp();
    py += 2 * sy;
    orow += 2LL * P.out_stride;
  }
}

// ---- sparse colour conversion (the pipeline's JPEG ingest).  The pipeline never shows the decoded frame to anybody: the letterbox
// reads the two source rows of every output row, the face and eye warps read the taps of their ROIs.  Only those pixels are
// converted: the rows of a list (jpeg_color_rows_kernel, before the detector) and the row spans roi_row_span gives for a warp's
// source quadrilateral (jpeg_color_roi_kernel, once the ROIs are known) -- the same spans roi_fill_kernel stages for zero-copy host
// frames, so the warps' arithmetic and their coverage argument are unchanged.  The pixel arithmetic is jpeg_color_kernel's.
// One warp: pixels [xa, xb) of row y (xa a multiple of 16: 48 bytes, so every 16-byte store is aligned; xb <= width).
__device__ __forceinline__ void color_row_span(const ColorShared& P, int y, int xa, int xb, uint32_t* stage, int lane) {
  const uint8_t *yrow, *nb, *nr, *fb, *fr;
  color_row_ptrs(P, y, &yrow, &nb, &nr, &fb, &fr);
  uint8_t* orow = P.out + (long long)y * P.out_stride;
  for (int xw = xa; xw < xb; xw += 128) {
    const int x0 = xw + 4 * lane;
    uint32_t rgb[3] = {0, 0, 0};
    if (x0 < xb) color_px4(P, yrow, nb, nr, fb, fr, x0, rgb);
    stage[3 * lane] = rgb[0]; stage[3 * lane + 1] = rgb[1]; stage[3 * lane + 2] = rgb[2];
    __syncwarp();
    const int row_bytes = 3 * min(128, xb - xw);
    if (lane < 24 && 16 * lane < row_bytes) {
      uint8_t* o = orow + 3LL * xw + 16 * lane;
      if (16 * lane + 16 <= row_bytes) *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(&stage[4 * lane]);
      else {
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(&stage[4 * lane]);
        for (int i = 0; 16 * lane + i < row_bytes; ++i) o[i] = sb[i];
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) jpeg_color_rows_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                                              uint8_t* __restrict__ out, const int* __restrict__ rows, int nrows) {
  __shared__ ColorShared P;
  __shared__ __align__(16) uint32_t s_rgb[8][96];
  if (threadIdx.x == 0) color_shared_fill(P, descs[blockIdx.y], planes, out);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 8 + warp;
  if (idx >= nrows) return;
  const int y = rows[idx];
  if (y < 0 || y >= P.height) return;
  color_row_span(P, y, 0, P.width, s_rgb[warp], lane);
}

// CTA k of a slot's gridDim.x CTAs takes the rows y0 + 8 * k + warp (+ 8 * gridDim.x, ...): 8 CTAs for a face, 2 for an eye (the
// box arithmetic is one thread's serial f64 work per CTA)

__global__ void __launch_bounds__(256) jpeg_color_roi_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                                             uint8_t* __restrict__ out, const I2TParams* __restrict__ params, int n,
                                                             const int* n_active, int n_images, const I2TParams* __restrict__ parents,
                                                             const uint8_t* __restrict__ rows_done) {
  if (n_active) n = min(n, *n_active);
  const int slot = blockIdx.y;
  if (slot >= n) return;
  __shared__ ColorShared P;
  __shared__ SrcBox s_box;
  __shared__ int s_margin;
  __shared__ __align__(16) uint32_t s_rgb[8][96];
  if (threadIdx.x == 0) {
    const I2TParams& Q = params[slot];
    int m = 0;
    SrcBox b = roi_stage_box(Q, 0, true, &m);
    if (Q.frame < 0 || Q.frame >= n_images) { b.x1 = -1; b.y1 = -1; }
    else if (parents && b.x1 >= b.x0) {
      // an eye slot whose source region lies inside what its face's pass converted (the test eye_split_kernel makes for zero-copy
      // frames) has nothing left to do
      int fm = 0;
      const SrcBox f = roi_stage_box(parents[slot >> 1], 0, true, &fm);
      if (parents[slot >> 1].frame == Q.frame && roi_stage_covers(f, warp_src_box(Q), Q.src_w, Q.src_h)) { b.x1 = -1; b.y1 = -1; }
    }
    if (b.x1 >= b.x0) color_shared_fill(P, descs[Q.frame], planes, out);
    s_box = b; s_margin = m;
  }
  __syncthreads();
  if (s_box.x1 < s_box.x0 || s_box.y1 < s_box.y0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r1 = min(s_box.y1, P.height - 1);
  // The warp's rows are y0 + 8 * k + warp + j * step.  Their spans (roi_row_span: a few hundred f64 instructions, more than the
  // pixels of a face-sized row cost) are computed 32 rows at a time, one row per lane, and handed round by shuffle.
  const int rstep = 8 * (int)gridDim.x;
  for (int rb = s_box.y0 + 8 * (int)blockIdx.x + warp; rb <= r1; rb += 32 * rstep) {
    const int rl = rb + lane * rstep;
    int x0l = 0, x1l = -1;
    if (rl <= r1 && !(rows_done && rows_done[rl])) {     // (rows_done: a whole row the rows pass converted)
      int a, b;
      if (roi_row_span(s_box, rl, s_margin, &a, &b)) { x0l = a; x1l = b; }
    }
    for (int j = 0; j < 32; ++j) {
      const int r = rb + j * rstep;
      if (r > r1) break;
      const int x0 = __shfl_sync(0xffffffffu, x0l, j), x1 = __shfl_sync(0xffffffffu, x1l, j);
      if (x1 < x0) continue;
      const int xa = max(x0, 0) & ~15, xb = min((x1 + 16) & ~15, P.width);
      if (xb > xa) color_row_span(P, r, xa, xb, s_rgb[warp], lane);
    }
  }
}

}  // namespace

// debugging aid: phase timestamps (ns) of the last entropy launch's CTA 0
cudaError_t jpeg_debug_phases(long long out[8]) { return cudaMemcpyFromSymbol(out, g_jpeg_phase, 8 * sizeof(long long)); }

int jpeg_scan_tiles(long long raw_off, int raw_len) { return (int)(((raw_off & 15) + raw_len + kScanTile - 1) / kScanTile); }

size_t jpeg_entropy_smem_bytes(int max_windows) {
  return ((sizeof(EntropyShared) + 15) & ~size_t(15)) + ((size_t)6 << (kLookBits + 1)) + (size_t)max_windows * (8 + 4 + 4 + 4 + 12);
}

cudaError_t launch_jpeg_entropy(const JpegImageDesc* descs, int n, const JpegHuff* tabs, const uint8_t* bytes, uint8_t* clean, int16_t* coef,
                                int* iv, int* status, int* tile_info, int max_tiles, int* scan_len, int max_windows, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  opt_in_smem_once(g_entropy_optin, jpeg_entropy_kernel, (int)jpeg_entropy_smem_bytes(kJpegMaxWindows));
  jpeg_scan_count_kernel<<<dim3((unsigned)max_tiles, (unsigned)n), kScanThreads, 0, s>>>(descs, bytes, tile_info, max_tiles);
  count_launch();
  jpeg_scan_compact_kernel<<<dim3((unsigned)max_tiles, (unsigned)n), kScanThreads, 0, s>>>(descs, bytes, tile_info, max_tiles, clean, iv, scan_len);
  count_launch();
  jpeg_entropy_kernel<<<n, kJpegEntropyThreads, jpeg_entropy_smem_bytes(max_windows), s>>>(descs, tabs, scan_len, clean, coef, iv, status);
  count_launch();
  return cudaGetLastError();
}

bool jpeg_idct_clears_coef() {
  static const bool per_block = getenv("FDL_JPEG_IDCT") ? atoi(getenv("FDL_JPEG_IDCT")) != 0 : true;
  return per_block;
}

cudaError_t launch_jpeg_idct(const JpegImageDesc* descs, int n, int max_quads, int16_t* coef, uint8_t* planes, cudaStream_t s) {
  if (n <= 0 || max_quads <= 0) return cudaSuccess;
  const bool per_block = jpeg_idct_clears_coef();
  if (per_block) jpeg_idct_block_kernel<<<dim3((unsigned)((4 * max_quads + 127) / 128), (unsigned)n), 128, 0, s>>>(descs, coef, planes);
  else jpeg_idct_kernel<<<dim3((unsigned)((max_quads + 7) / 8), (unsigned)n), 256, 0, s>>>(descs, coef, planes);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_jpeg_color(const JpegImageDesc* descs, int n, int max_w, int max_h, int flags, const uint8_t* planes, uint8_t* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (flags & 1) { jpeg_color_kernel<<<dim3((unsigned)((max_w + 255) / 256), (unsigned)((max_h + 8 * kColorPairs - 1) / (8 * kColorPairs)), (unsigned)n), 256, 0, s>>>(descs, planes, out); count_launch(); }
  if (flags & 2) { jpeg_color_generic_kernel<<<dim3((unsigned)((max_w + 1023) / 1024), (unsigned)max_h, (unsigned)n), 256, 0, s>>>(descs, planes, out); count_launch(); }
  return cudaGetLastError();
}

cudaError_t launch_jpeg_color_rows(const JpegImageDesc* descs, int n, const int* rows, int nrows, const uint8_t* planes, uint8_t* out, cudaStream_t s) {
  if (n <= 0 || nrows <= 0) return cudaSuccess;
  jpeg_color_rows_kernel<<<dim3((unsigned)((nrows + 7) / 8), (unsigned)n), 256, 0, s>>>(descs, planes, out, rows, nrows);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_jpeg_color_roi(const JpegImageDesc* descs, int n_images, const I2TParams* params, int n, const int* n_active,
                                  const I2TParams* parents, const uint8_t* rows_done, const uint8_t* planes, uint8_t* out, cudaStream_t s) {
  if (n <= 0 || n_images <= 0) return cudaSuccess;
  jpeg_color_roi_kernel<<<dim3(parents ? 2 : 8, (unsigned)n), 256, 0, s>>>(descs, planes, out, params, n, n_active, n_images, parents, rows_done);
  count_launch();
  return cudaGetLastError();
}

}  // namespace fdl
