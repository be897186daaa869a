// jpeg_decode.h -- host side of the device JPEG decoder (see jpeg_device.h): parses the headers of a batch of JPEG files
// (jpeg_parse.h: the oracle's acceptance rules), lays the batch out in grow-only work buffers, stages bytes + descriptors +
// Huffman tables and enqueues the three kernels.  One JpegDecoder per stream of batches (a pipeline lane, a decoder handle):
// a batch's buffers are reused by the next plan() call, so the previous batch must have been waited for.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "device_util.h"
#include "glue_math.h"
#include "jpeg_device.h"

namespace fdl {

class JpegDecoder {
 public:
  // Parse + lay out.  expect_w / expect_h > 0: every image must have exactly that size.  Returns FDL_OK or an error code
  // (message set); nothing is enqueued.
  int plan(const uint8_t* const* data, const size_t* len, int n, int expect_w, int expect_h);
  int count() const { return n_; }
  int width(int i) const { return h_descs_.p[i].width; }
  int height(int i) const { return h_descs_.p[i].height; }
  int components(int i) const { return h_descs_.p[i].ncomp; }
  size_t compressed_bytes() const { return total_bytes_; }
  // Where image i goes in the output buffer handed to enqueue() (bytes; rows of out_stride bytes).
  void set_output(int i, long long out_off, int out_stride) { h_descs_.p[i].out_off = out_off; h_descs_.p[i].out_stride = out_stride; }
  // H2D of the compressed bytes / descriptors / tables, coefficient clear, entropy + IDCT + colour kernels, status D2H -- all on `s`.
  // `rows` (device list of `nrows` frame rows) != null asks for the sparse conversion: only those rows of every image are converted
  // here, the caller converts the regions its warps read with color_roi() once it knows them.  *sparse tells whether that was
  // possible (every image on the colour fast path); otherwise the whole frames were converted as usual.
  // `rows_done`: the same rows as a byte mask over the frame rows (color_roi skips them).
  int enqueue(uint8_t* out_device, cudaStream_t s, const int* rows = nullptr, int nrows = 0, const uint8_t* rows_done = nullptr, bool* sparse = nullptr);
  // After a sparse enqueue: convert the source regions of `n` image_to_tensor slots (device parameters; *n_active bounds n).
  // `parents`: the slots (one per two of `params`: a face and its eyes) whose regions an earlier call converted.
  cudaError_t color_roi(const I2TParams* params, int n, const int* n_active, const I2TParams* parents, cudaStream_t s);
  // After `s` has been waited for: FDL_OK, or FDL_ERR_INVALID naming the first image whose entropy-coded data was inconsistent.
  int check_status();

 private:
  int intern_table(const uint8_t* dht);
  uint8_t* out_device_ = nullptr;
  const uint8_t* rows_done_ = nullptr;
  int n_ = 0;
  size_t total_bytes_ = 0, span_bytes_ = 0, clean_bytes_ = 0, coef_elems_ = 0, plane_bytes_ = 0, iv_entries_ = 0;
  int max_windows_ = 1, max_quads_ = 0, max_w_ = 0, max_h_ = 0, max_tiles_ = 1;
  const uint8_t* direct_src_ = nullptr;     // pinned caller memory copied without staging (one span), or null
  std::vector<std::vector<uint8_t>> dht_blobs_;
  DevBuf<uint8_t> d_bytes_, d_clean_, d_planes_;
  DevBuf<int16_t> d_coef_;
  DevBuf<int> d_iv_, d_status_, d_tile_info_, d_scan_len_;
  DevBuf<JpegImageDesc> d_descs_;
  DevBuf<JpegHuff> d_tabs_;
  PinBuf<uint8_t> h_bytes_;
  PinBuf<JpegImageDesc> h_descs_;
  PinBuf<JpegHuff> h_tabs_;
  PinBuf<int> h_status_;
};

}  // namespace fdl
