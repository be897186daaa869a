// jpeg_decode.cu -- see jpeg_decode.h.
#include "jpeg_decode.h"

#include <cstdlib>
#include <cstring>

#include "fdl_status.h"
#include "jpeg_parse.h"

namespace fdl {

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

int JpegDecoder::intern_table(const uint8_t* dht) {
  for (size_t i = 0; i < dht_blobs_.size(); ++i)
    if (std::memcmp(dht_blobs_[i].data(), dht, 16 + 256) == 0) return (int)i;
  dht_blobs_.emplace_back(dht, dht + 16 + 256);
  return (int)dht_blobs_.size() - 1;
}

int JpegDecoder::plan(const uint8_t* const* data, const size_t* len, int n, int expect_w, int expect_h) {
  if (!data || !len || n <= 0) return set_error(FDL_ERR_INVALID, "no JPEG data given");
  n_ = 0;
  FDL_CUDA_TRY(h_descs_.reserve((size_t)n));
  FDL_CUDA_TRY(h_status_.reserve((size_t)n));
  dht_blobs_.clear();
  total_bytes_ = clean_bytes_ = coef_elems_ = plane_bytes_ = iv_entries_ = 0;
  max_windows_ = 1; max_quads_ = 0; max_w_ = max_h_ = 0; max_tiles_ = 1;

  // Where the compressed bytes come from: if the caller's buffers are pinned and lie close together in ascending order (one
  // arena of encoded frames), the H2D copy reads them in place as one span; otherwise they are packed into a pinned staging buffer.
  const uint8_t* lo = data[0];
  const uint8_t* hi = data[0];
  bool ascending = true;
  size_t sum = 0;
  for (int i = 0; i < n; ++i) {
    if (!data[i] || len[i] < 4) return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + ": not a JPEG");
    if (len[i] >= ((size_t)1 << 28)) return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + ": larger than 256 MiB");
    if (i && data[i] < data[i - 1] + len[i - 1]) ascending = false;
    if (data[i] < lo) lo = data[i];
    if (data[i] + len[i] > hi) hi = data[i] + len[i];
    sum += len[i];
  }
  direct_src_ = nullptr;
  if (ascending && (size_t)(hi - lo) <= sum + 256 * (size_t)n) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, lo) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      cudaPointerAttributes at2;
      if (cudaPointerGetAttributes(&at2, hi - 1) == cudaSuccess && at2.type == cudaMemoryTypeHost) direct_src_ = lo;
    }
    cudaGetLastError();
  }
  std::vector<size_t> byte_off((size_t)n);
  if (direct_src_) {
    for (int i = 0; i < n; ++i) byte_off[(size_t)i] = (size_t)(data[i] - lo);
    span_bytes_ = (size_t)(hi - lo);
  } else {
    size_t o = 0;
    for (int i = 0; i < n; ++i) { byte_off[(size_t)i] = o; o += align_up(len[i], 16); }
    span_bytes_ = o;
    FDL_CUDA_TRY(h_bytes_.reserve(span_bytes_ + 64));
    for (int i = 0; i < n; ++i) std::memcpy(h_bytes_.p + byte_off[(size_t)i], data[i], len[i]);
  }
  total_bytes_ = sum;

  for (int i = 0; i < n; ++i) {
    JpegHeader hd;
    std::string msg;
    if (!jpeg_parse_header(data[i], len[i], &hd, &msg)) return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + ": " + msg);
    if ((expect_w > 0 && hd.width != expect_w) || (expect_h > 0 && hd.height != expect_h))
      return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + ": " + std::to_string(hd.width) + "x" + std::to_string(hd.height) +
                                            " differs from the expected frame size");
    JpegImageDesc& d = h_descs_.p[i];
    std::memset(&d, 0, sizeof(d));
    d.raw_off = (long long)(byte_off[(size_t)i] + hd.scan_offset);
    d.raw_len = (int)(len[i] - hd.scan_offset);
    d.width = hd.width; d.height = hd.height; d.ncomp = hd.ncomp;
    d.hmax = hd.hmax; d.vmax = hd.vmax; d.mcux = hd.mcus_x; d.mcuy = hd.mcus_y;
    int bpm = 0, quads = 0;
    for (int c = 0; c < hd.ncomp; ++c) {
      const JpegComponent& k = hd.comp[c];
      d.hs[c] = k.h; d.vs[c] = k.v;
      for (int by = 0; by < k.v; ++by)
        for (int bx = 0; bx < k.h; ++bx) {
          if (bpm >= kJpegMaxBlocksPerMcu) return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + ": more than 10 blocks per MCU");
          d.blk_comp[bpm] = c; d.blk_bx[bpm] = bx; d.blk_by[bpm] = by; ++bpm;
        }
      d.tab_dc[c] = intern_table(hd.dht[0][k.td]);
      d.tab_ac[c] = intern_table(hd.dht[1][k.ta]);
      std::memcpy(d.quant[c], hd.quant[k.tq], sizeof(d.quant[c]));
      d.bcols[c] = hd.mcus_x * k.h; d.brows[c] = hd.mcus_y * k.v;
      d.cw[c] = (hd.width * k.h + hd.hmax - 1) / hd.hmax; d.ch[c] = (hd.height * k.v + hd.vmax - 1) / hd.vmax;
      d.coef_off[c] = (long long)coef_elems_;
      coef_elems_ += align_up((size_t)d.bcols[c] * d.brows[c] * 64, 8);
      d.plane_off[c] = (long long)plane_bytes_;
      plane_bytes_ += align_up((size_t)d.bcols[c] * 8 * d.brows[c] * 8, 16);
      quads += d.brows[c] * ((d.bcols[c] + 3) / 4);
    }
    d.bpm = bpm;
    d.restart_interval = hd.restart_interval;
    const long long mcus = (long long)hd.mcus_x * hd.mcus_y;
    d.n_intervals = hd.restart_interval > 0 ? (int)((mcus + hd.restart_interval - 1) / hd.restart_interval) : 0;
    d.iv_off = (long long)iv_entries_;
    iv_entries_ += (size_t)d.n_intervals;
    d.clean_off = (long long)clean_bytes_;
    clean_bytes_ += align_up((size_t)d.raw_len + 64, 128);     // 128-byte aligned: a 1024-bit window is one L1 line
    const long long bits = (long long)d.raw_len * 8;
    long long wb = (bits + kJpegMaxWindows - 1) / kJpegMaxWindows;
    wb = (wb + 31) / 32 * 32;
    static const int min_bits = [] { const char* e = getenv("FDL_JPEG_WINDOW_BITS"); const int v = e ? atoi(e) : 0; return v >= 64 ? v / 32 * 32 : kJpegMinWindowBits; }();
    d.window_bits = (int)(wb < min_bits ? min_bits : wb);
    d.nwin_cap = (int)((bits + d.window_bits - 1) / d.window_bits);
    if (d.nwin_cap < 1) d.nwin_cap = 1;
    if (d.restart_interval == 0 && d.nwin_cap > max_windows_) max_windows_ = d.nwin_cap;
    if (quads > max_quads_) max_quads_ = quads;
    const int tiles = jpeg_scan_tiles(d.raw_off, d.raw_len);
    if (tiles > max_tiles_) max_tiles_ = tiles;
    if (d.width > max_w_) max_w_ = d.width;
    if (d.height > max_h_) max_h_ = d.height;
    d.out_off = 0; d.out_stride = d.width * 3;
  }
  FDL_CUDA_TRY(h_tabs_.reserve(dht_blobs_.size()));
  for (size_t t = 0; t < dht_blobs_.size(); ++t) jpeg_huff_build(dht_blobs_[t].data(), dht_blobs_[t].data() + 16, &h_tabs_.p[t]);
  n_ = n;
  return FDL_OK;
}

int JpegDecoder::enqueue(uint8_t* out_device, cudaStream_t s, const int* rows, int nrows, const uint8_t* rows_done, bool* sparse) {
  if (n_ <= 0) return set_error(FDL_ERR_INVALID, "no planned JPEG batch");
  FDL_CUDA_TRY(d_bytes_.reserve(span_bytes_ + 64));
  FDL_CUDA_TRY(d_clean_.reserve(clean_bytes_ + 256));
  const int16_t* coef_before = d_coef_.p;
  FDL_CUDA_TRY(d_coef_.reserve(coef_elems_));
  const bool coef_is_new = d_coef_.p != coef_before;
  FDL_CUDA_TRY(d_planes_.reserve(plane_bytes_ + 64));      // the colour kernel's word loads may run a few bytes past the last row
  FDL_CUDA_TRY(d_iv_.reserve(iv_entries_ + 1));
  FDL_CUDA_TRY(d_status_.reserve((size_t)n_));
  FDL_CUDA_TRY(d_tile_info_.reserve((size_t)n_ * max_tiles_ * 3));
  FDL_CUDA_TRY(d_scan_len_.reserve((size_t)n_ * 2));
  FDL_CUDA_TRY(d_descs_.reserve((size_t)n_));
  FDL_CUDA_TRY(d_tabs_.reserve(dht_blobs_.size()));
  // which colour kernel takes which image (the output placement is known only now)
  int color_flags = 0;
  for (int i = 0; i < n_; ++i) {
    JpegImageDesc& d = h_descs_.p[i];
    const bool same = d.ncomp == 3 && d.hs[0] == d.hmax && d.vs[0] == d.vmax && d.hs[1] == d.hs[2] && d.vs[1] == d.vs[2] && d.hmax / d.hs[1] == 2 &&
                      d.cw[1] > 2 && d.cw[1] == d.cw[2] && d.ch[1] == d.ch[2];
    const bool aligned = d.out_stride % 16 == 0 && d.out_off % 16 == 0 && (reinterpret_cast<uintptr_t>(out_device) & 15) == 0;
    d.color_fast = same && aligned ? 1 : 0;
    color_flags |= d.color_fast ? 1 : 2;
  }
  FDL_CUDA_TRY(cudaMemcpyAsync(d_bytes_.p, direct_src_ ? direct_src_ : h_bytes_.p, span_bytes_, cudaMemcpyHostToDevice, s));
  FDL_CUDA_TRY(cudaMemcpyAsync(d_descs_.p, h_descs_.p, (size_t)n_ * sizeof(JpegImageDesc), cudaMemcpyHostToDevice, s));
  FDL_CUDA_TRY(cudaMemcpyAsync(d_tabs_.p, h_tabs_.p, dht_blobs_.size() * sizeof(JpegHuff), cudaMemcpyHostToDevice, s));
  // the entropy stage writes non-zero coefficients into a zeroed buffer; the IDCT kernel leaves it zeroed again
  if (coef_is_new) FDL_CUDA_TRY(cudaMemsetAsync(d_coef_.p, 0, d_coef_.cap * sizeof(int16_t), s));
  else if (!jpeg_idct_clears_coef()) FDL_CUDA_TRY(cudaMemsetAsync(d_coef_.p, 0, coef_elems_ * sizeof(int16_t), s));
  FDL_CUDA_TRY(launch_jpeg_entropy(d_descs_.p, n_, d_tabs_.p, d_bytes_.p, d_clean_.p, d_coef_.p, d_iv_.p, d_status_.p, d_tile_info_.p, max_tiles_, d_scan_len_.p, max_windows_, s));
  FDL_CUDA_TRY(launch_jpeg_idct(d_descs_.p, n_, max_quads_, d_coef_.p, d_planes_.p, s));
  const bool do_sparse = rows && nrows > 0 && color_flags == 1;
  if (sparse) *sparse = do_sparse;
  out_device_ = out_device;
  rows_done_ = do_sparse ? rows_done : nullptr;
  if (do_sparse) FDL_CUDA_TRY(launch_jpeg_color_rows(d_descs_.p, n_, rows, nrows, d_planes_.p, out_device, s));
  else FDL_CUDA_TRY(launch_jpeg_color(d_descs_.p, n_, max_w_, max_h_, color_flags, d_planes_.p, out_device, s));
  FDL_CUDA_TRY(cudaMemcpyAsync(h_status_.p, d_status_.p, (size_t)n_ * sizeof(int), cudaMemcpyDeviceToHost, s));
  return FDL_OK;
}

cudaError_t JpegDecoder::color_roi(const I2TParams* params, int n, const int* n_active, const I2TParams* parents, cudaStream_t s) {
  return launch_jpeg_color_roi(d_descs_.p, n_, params, n, n_active, parents, rows_done_, d_planes_.p, out_device_, s);
}

int JpegDecoder::check_status() {
  for (int i = 0; i < n_; ++i) {
    if (h_status_.p[i] == JPEG_OK) continue;
    return set_error(FDL_ERR_INVALID, "image " + std::to_string(i) + (h_status_.p[i] == JPEG_ERR_RESTARTS
                                                                          ? ": fewer restart markers than the restart interval demands"
                                                                          : ": premature end of the entropy-coded data"));
  }
  return FDL_OK;
}

}  // namespace fdl
