// jpeg_kernels.cu -- see jpeg_device.h.  The three kernels of the device JPEG decoder (frame ingest, utils.rs:8-21).
#include "jpeg_device.h"

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
struct EntropyShared {
  JpegHuff tabs[6];          // [2 * component + {0: DC, 1: AC}]
  JpegImageDesc d;
  int warp_part[32];
  int scan_total;
  int term;                  // first byte (relative to the aligned start) of the marker that ends the scan
  uint8_t zz[64];
};

// 32 bits of the clean scan starting at bit `pos` (MSB first).  The scan is followed by 32 zero bytes.
__device__ __forceinline__ uint32_t peek32(const uint32_t* w, uint32_t pos) {
  const uint32_t i = pos >> 5;
  const uint32_t a = __byte_perm(w[i], 0, 0x0123), b = __byte_perm(w[i + 1], 0, 0x0123);
  return __funnelshift_l(b, a, pos & 31);
}

struct DecState { uint32_t pos; int b, k; };

// Decodes symbols from `st` while st.pos < end and fewer than max_blocks blocks have been completed -- jpeg_sync_step
// (jpeg_math.h) with a 32-bit peek: a Huffman code (<= 16 bits) and its extra bits (<= 16) always fit.
//   WRITE = false: dc[c] accumulates the DC differences of component c (the window's contribution to the predictors).
//   WRITE = true:  dc[c] are the running predictors; coefficients go to their blocks, the first being block `b` of MCU `m`.
// Returns the number of blocks completed.
template <bool WRITE>
__device__ __forceinline__ int jpeg_run(const uint32_t* w, const EntropyShared& S, DecState& st, uint32_t end, int max_blocks, int dc[3],
                                        int16_t* coef, int m) {
  const JpegImageDesc& d = S.d;
  uint32_t pos = st.pos;
  int b = st.b, k = st.k, nb = 0;
  int mx = 0, my = 0;
  int16_t* blk = nullptr;
  if (WRITE) {
    my = m / d.mcux; mx = m - my * d.mcux;
    const int c = d.blk_comp[b];
    blk = coef + d.coef_off[c] + ((long long)(my * d.vs[c] + d.blk_by[b]) * d.bcols[c] + mx * d.hs[c] + d.blk_bx[b]) * 64;
  }
  while (pos < end && nb < max_blocks) {
    const int c = d.blk_comp[b];
    const JpegHuff& t = S.tabs[2 * c + (k ? 1 : 0)];
    const uint32_t v = peek32(w, pos);
    int len, sym;
    const int look = t.look[v >> 23];
    if (look) { len = look >> 8; sym = look & 255; }
    else {
      len = 10;
      int code = (int)(v >> 22);
      while (len <= 16 && code > t.maxcode[len]) { ++len; code = (int)(v >> (32 - len)); }
      if (len > 16) { len = 16; sym = 0; } else sym = t.huffval[(code + t.valoffset[len]) & 255];
    }
    bool done = false;
    if (k == 0) {                                     // DC: the symbol is the number of extra bits
      const int sz = sym > 16 ? 16 : sym;
      int diff = 0;
      if (sz) diff = jpeg_extend((int)((v << len) >> (32 - sz)), sz);
      pos += len + sz;
      dc[c] += diff;
      if (WRITE) blk[0] = (int16_t)dc[c];
      k = 1;
    } else {
      const int r = sym >> 4, s = sym & 15;
      pos += len;
      if (s == 0) {
        if (r != 15) done = true;                     // EOB
        else { k += 16; done = k > 63; }              // ZRL
      } else {
        k += r;
        if (k > 63) done = true;                      // corrupt run: the block ends, as in the sequential decoder
        else {
          if (WRITE) blk[S.zz[k]] = (int16_t)jpeg_extend((int)((v << len) >> (32 - s)), s);
          pos += s;
          done = ++k > 63;
        }
      }
    }
    if (done) {
      k = 0; ++nb;
      if (++b == d.bpm) { b = 0; if (WRITE) { if (++mx == d.mcux) { mx = 0; ++my; } } }
      if (WRITE) {
        const int c2 = d.blk_comp[b];
        blk = coef + d.coef_off[c2] + ((long long)(my * d.vs[c2] + d.blk_by[b]) * d.bcols[c2] + mx * d.hs[c2] + d.blk_bx[b]) * 64;
      }
    }
  }
  st.pos = pos; st.b = b; st.k = k;
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

__global__ void __launch_bounds__(kJpegEntropyThreads, 1)
jpeg_entropy_kernel(const JpegImageDesc* __restrict__ descs, const JpegHuff* __restrict__ tabs, const uint8_t* __restrict__ bytes,
                    uint8_t* clean_all, int16_t* coef, int* iv_all, int* status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EntropyShared& S = *reinterpret_cast<EntropyShared*>(smem_raw);
  const int tid = threadIdx.x, T = blockDim.x, img = blockIdx.x;
  // ---- descriptor, tables
  {
    const int* src = reinterpret_cast<const int*>(descs + img);
    int* dst = reinterpret_cast<int*>(&S.d);
    for (int i = tid; i < (int)(sizeof(JpegImageDesc) / 4); i += T) dst[i] = src[i];
  }
  __syncthreads();
  const JpegImageDesc& d = S.d;
  for (int c = 0; c < d.ncomp; ++c)
    for (int a = 0; a < 2; ++a) {
      const int* src = reinterpret_cast<const int*>(tabs + (a ? d.tab_ac[c] : d.tab_dc[c]));
      int* dst = reinterpret_cast<int*>(&S.tabs[2 * c + a]);
      for (int i = tid; i < (int)(sizeof(JpegHuff) / 4); i += T) dst[i] = src[i];
    }
  if (tid < 64) {
    const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    S.zz[tid] = zz[tid];
  }
  if (tid == 0) S.term = INT_MAX;
  __syncthreads();

  uint8_t* clean = clean_all + d.clean_off;
  int* ivs = iv_all + d.iv_off;

  // ---- 1. the scan without stuffing and restart markers (a stream compaction): FF00 -> FF, FFDn dropped and remembered as an interval
  // start, FFFF fill dropped, any other marker ends the scan.  16 bytes per thread and tile.
  const long long lo_abs = d.raw_off, hi_abs = d.raw_off + d.raw_len, a0 = lo_abs & ~15LL;
  int run_bytes = 0, run_rst = 0;
  for (long long tile = a0; tile < hi_abs; tile += (long long)T * 16) {
    const long long abs = tile + (long long)tid * 16;
    uint32_t q[4] = {0, 0, 0, 0};
    int prev = 0, next = 0xD9;
    if (abs < hi_abs) {
      const uint4 u = *reinterpret_cast<const uint4*>(bytes + abs);
      q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w;
      if (abs - 1 >= lo_abs) prev = bytes[abs - 1];
      if (abs + 16 < hi_abs) next = bytes[abs + 16];
    }
    uint32_t keep = 0, rst = 0;
    int my_term = INT_MAX;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = (q[j >> 2] >> (8 * (j & 3))) & 255;
      const int nx = j < 15 ? (int)((q[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 255) : next;
      const long long pa = abs + j;
      if (pa >= lo_abs && pa < hi_abs) {
        const int nxe = pa + 1 < hi_abs ? nx : 0xD9;
        if (c == 0xFF) {
          if (nxe == 0) keep |= 1u << j;
          else if (nxe >= 0xD0 && nxe <= 0xD7) rst |= 1u << j;
          else if (nxe != 0xFF && my_term == INT_MAX) my_term = (int)(pa - a0);
        } else if (!(prev == 0xFF && (c == 0 || (c >= 0xD0 && c <= 0xD7)))) {
          keep |= 1u << j;
        }
      }
      prev = c;
    }
    if (my_term != INT_MAX) atomicMin(&S.term, my_term);
    __syncthreads();
    const int term = S.term;
    if (term != INT_MAX) {
      const long long first_dead = a0 + term - abs;        // bytes of this thread at index >= first_dead lie behind the end of the scan
      const uint32_t alive = first_dead >= 16 ? 0xFFFFu : (first_dead <= 0 ? 0u : ((1u << (int)first_dead) - 1u));
      keep &= alive; rst &= alive;
    }
    int total;
    const int v = __popc(keep) | (__popc(rst) << 16);
    const int excl = block_excl_scan(v, S, &total);
    int o = run_bytes + (excl & 0xFFFF), r = run_rst + (excl >> 16);
    if (keep | rst) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (keep & (1u << j)) clean[o++] = (uint8_t)((q[j >> 2] >> (8 * (j & 3))) & 255);
        else if (rst & (1u << j)) { if (++r < d.n_intervals) ivs[r] = o; }
      }
    }
    run_bytes += total & 0xFFFF; run_rst += total >> 16;
    if (term != INT_MAX) break;
  }
  if (tid < 32) clean[run_bytes + tid] = 0;
  if (tid == 0 && d.n_intervals > 0) ivs[0] = 0;
  __syncthreads();
  const uint32_t* w = reinterpret_cast<const uint32_t*>(clean);
  const uint32_t nbits = (uint32_t)run_bytes * 8u;
  const int total_blocks = d.mcux * d.mcuy * d.bpm;

  // ---- 2a. restart intervals: byte-aligned, predictors reset, position known from the interval index: one thread each
  if (d.restart_interval > 0) {
    const int mcus = d.mcux * d.mcuy;
    const bool ok = run_rst + 1 >= d.n_intervals;
    if (ok) {
      for (int iv = tid; iv < d.n_intervals; iv += T) {
        DecState st = {(uint32_t)ivs[iv] * 8u, 0, 0};
        const int m0 = iv * d.restart_interval, m1 = min(m0 + d.restart_interval, mcus);
        int pred[3] = {0, 0, 0};
        jpeg_run<true>(w, S, st, nbits, (m1 - m0) * d.bpm, pred, coef, m0);
      }
    }
    if (tid == 0) status[img] = ok ? JPEG_OK : JPEG_ERR_RESTARTS;
    return;
  }

  // ---- 2b. no restart markers: self-synchronising windows (Weissenberger & Schmidt, ICPP 2018; jpeg_math.h)
  const int WB = d.window_bits;
  const int nwin = min((int)((nbits + (uint32_t)WB - 1u) / (uint32_t)WB), d.nwin_cap);
  unsigned char* dyn = smem_raw + ((sizeof(EntropyShared) + 15) & ~size_t(15));
  unsigned long long* s_exit = reinterpret_cast<unsigned long long*>(dyn);          // pos | (b << 8 | k) << 32 | nb << 48
  uint32_t* s_epos = reinterpret_cast<uint32_t*>(s_exit + d.nwin_cap);               // entry state
  uint32_t* s_ebk = s_epos + d.nwin_cap;
  int* s_dc = reinterpret_cast<int*>(s_ebk + d.nwin_cap);                            // [3][nwin_cap] sums of DC differences
  const int per = (nwin + T - 1) / T;
  const int wlo = min(tid * per, nwin), whi = min(wlo + per, nwin);
  auto window_end = [&](int i) { const unsigned long long e = (unsigned long long)(i + 1) * (unsigned)WB; return e < nbits ? (uint32_t)e : nbits; };
  auto decode_window = [&](int i, DecState st) {
    s_epos[i] = st.pos; s_ebk[i] = (uint32_t)(st.b << 8 | st.k);
    int dc[3] = {0, 0, 0};
    const int nb = jpeg_run<false>(w, S, st, window_end(i), INT_MAX, dc, nullptr, 0);
    s_exit[i] = (unsigned long long)st.pos | ((unsigned long long)(st.b << 8 | st.k) << 32) | ((unsigned long long)nb << 48);
    s_dc[i] = dc[0]; s_dc[d.nwin_cap + i] = dc[1]; s_dc[2 * d.nwin_cap + i] = dc[2];
    return st;
  };
  // round 0: the first window of every thread starts from a guess (block 0, DC symbol next, at the window's first bit); the thread's
  // other windows continue from its own exit state
  {
    DecState st = {0, 0, 0};
    for (int i = wlo; i < whi; ++i) {
      if (i == wlo) st = DecState{(uint32_t)((unsigned long long)i * (unsigned)WB), 0, 0};
      st = decode_window(i, st);
    }
  }
  // hand-over rounds: a window whose predecessor's exit state differs from the entry state it was decoded from is decoded again;
  // a fixed point is the sequential decode.  Exit states are snapshotted between two barriers, so a round reads only the previous
  // round's values.
  __syncthreads();
  for (;;) {
    unsigned long long ex = 0;
    const bool have = wlo > 0 && wlo < whi;
    if (have) ex = s_exit[wlo - 1];
    __syncthreads();
    int changed = 0;
    if (have) {
      for (int i = wlo; i < whi; ++i) {
        if (i > wlo) ex = s_exit[i - 1];
        const uint32_t pos = (uint32_t)ex, bk = (uint32_t)(ex >> 32) & 0xFFFFu;
        if (pos == s_epos[i] && bk == s_ebk[i]) break;
        decode_window(i, DecState{pos, (int)(bk >> 8), (int)(bk & 255)});
        changed = 1;
      }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // block and DC prefix sums over the windows place every window's output
  int my_nb = 0, my_dc[3] = {0, 0, 0};
  for (int i = wlo; i < whi; ++i) {
    my_nb += (int)(s_exit[i] >> 48);
    my_dc[0] += s_dc[i]; my_dc[1] += s_dc[d.nwin_cap + i]; my_dc[2] += s_dc[2 * d.nwin_cap + i];
  }
  int tot_nb, tot;
  int g = block_excl_scan(my_nb, S, &tot_nb);
  int pred[3];
  pred[0] = block_excl_scan(my_dc[0], S, &tot);
  pred[1] = block_excl_scan(my_dc[1], S, &tot);
  pred[2] = block_excl_scan(my_dc[2], S, &tot);
  // output pass: the same decode from the (now true) entry states, writing coefficients and absolute DC values
  for (int i = wlo; i < whi && g < total_blocks; ++i) {
    DecState st = {s_epos[i], (int)(s_ebk[i] >> 8), (int)(s_ebk[i] & 255)};
    g += jpeg_run<true>(w, S, st, window_end(i), total_blocks - g, pred, coef, g / d.bpm);
  }
  if (tid == 0) status[img] = tot_nb >= total_blocks ? JPEG_OK : JPEG_ERR_BLOCKS;
}

// ------------------------------------------------------------------------------------------------ dequantise + IDCT
// One warp = four horizontally adjacent blocks of one component; lane (j, t) loads row t of block j (16 bytes), runs column t,
// then row t, and stores 8 samples: the four blocks' rows make whole 32-byte sectors of the plane.
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegImageDesc* __restrict__ descs, const int16_t* __restrict__ coef,
                                                        uint8_t* __restrict__ planes) {
  __shared__ int ws[8][4][8][9];
  __shared__ uint16_t s_quant[3][64];
  const JpegImageDesc& d = descs[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = lane >> 3, t = lane & 7;
  if (tid < 192) s_quant[tid >> 6][tid & 63] = d.quant[tid >> 6][tid & 63];
  __syncthreads();
  int q = blockIdx.x * 8 + warp, c = 0;
  for (; c < d.ncomp; ++c) {
    const int qc = d.brows[c] * ((d.bcols[c] + 3) >> 2);
    if (q < qc) break;
    q -= qc;
  }
  if (c >= d.ncomp) return;
  const int qpr = (d.bcols[c] + 3) >> 2, row = q / qpr, bx = (q - row * qpr) * 4 + j;
  const bool valid = bx < d.bcols[c];
  int (*m)[9] = ws[warp][j];
  if (valid) {
    const int4 u = *reinterpret_cast<const int4*>(coef + d.coef_off[c] + ((long long)row * d.bcols[c] + bx) * 64 + t * 8);
    const int v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m[t][2 * i] = (int)(int16_t)(v[i] & 0xFFFF) * (int)s_quant[c][t * 8 + 2 * i];
      m[t][2 * i + 1] = (v[i] >> 16) * (int)s_quant[c][t * 8 + 2 * i + 1];
    }
  }
  __syncwarp();
  int x[8];
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = m[i][t];
    jpeg_idct_1d(x, 1, 13, 13 - 2);                                // column t
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i][t] = x[i];
  }
  __syncwarp();
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = m[t][i];
    jpeg_idct_1d(x, 1, 13, 13 + 2 + 3);                            // row t
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lo |= (uint32_t)jpeg_range_limit(x[i]) << (8 * i);
      hi |= (uint32_t)jpeg_range_limit(x[4 + i]) << (8 * i);
    }
    *reinterpret_cast<uint2*>(planes + d.plane_off[c] + (long long)(row * 8 + t) * (d.bcols[c] * 8) + bx * 8) = make_uint2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------------ upsampling + colour conversion
// Four chroma samples for the pixels x0 .. x0+3 (x0 a multiple of 4) of row y: jdsample.c's fancy upsampling in gather form, the
// column sums shared between the pixels (jpeg_h2v2_fancy_at / jpeg_h2v1_fancy_at of jpeg_math.h give the same values one at a time).
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

__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                                         uint8_t* __restrict__ out) {
  const JpegImageDesc& d = descs[blockIdx.z];
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
  if (x0 + 3 < d.width && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
    uint32_t* o4 = reinterpret_cast<uint32_t*>(o);
#pragma unroll
    for (int i = 0; i < 3; ++i) o4[i] = px[4 * i] | (uint32_t)px[4 * i + 1] << 8 | (uint32_t)px[4 * i + 2] << 16 | (uint32_t)px[4 * i + 3] << 24;
  } else {
    const int n = min(4, d.width - x0) * 3;
    for (int i = 0; i < n; ++i) o[i] = px[i];
  }
}

std::atomic<unsigned long long> g_entropy_optin{0};

}  // namespace

size_t jpeg_entropy_smem_bytes(int max_windows) {
  return ((sizeof(EntropyShared) + 15) & ~size_t(15)) + (size_t)max_windows * (8 + 4 + 4 + 12);
}

cudaError_t launch_jpeg_entropy(const JpegImageDesc* descs, int n, const JpegHuff* tabs, const uint8_t* bytes, uint8_t* clean, int16_t* coef,
                                int* iv, int* status, int max_windows, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  opt_in_smem_once(g_entropy_optin, jpeg_entropy_kernel, (int)jpeg_entropy_smem_bytes(kJpegMaxWindows));
  jpeg_entropy_kernel<<<n, kJpegEntropyThreads, jpeg_entropy_smem_bytes(max_windows), s>>>(descs, tabs, bytes, clean, coef, iv, status);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_jpeg_idct(const JpegImageDesc* descs, int n, int max_quads, const int16_t* coef, uint8_t* planes, cudaStream_t s) {
  if (n <= 0 || max_quads <= 0) return cudaSuccess;
  jpeg_idct_kernel<<<dim3((unsigned)((max_quads + 7) / 8), (unsigned)n), 256, 0, s>>>(descs, coef, planes);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_jpeg_color(const JpegImageDesc* descs, int n, int max_w, int max_h, const uint8_t* planes, uint8_t* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  jpeg_color_kernel<<<dim3((unsigned)((max_w + 1023) / 1024), (unsigned)max_h, (unsigned)n), 256, 0, s>>>(descs, planes, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace fdl
