// jpeg_device.h -- device JPEG decoder of the frame-ingest row: what replaces
// `convert_image_to_mat` (reference src/face_detection_lite/utils.rs:8-21: imgcodecs::imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB),
// i.e. OpenCV's bundled libjpeg with JDCT_ISLOW and fancy upsampling) on the way into the pipeline.  The arithmetic is
// csrc/jpeg_math.h (pinned bit-exact against cv2.imdecode on the host); this header holds the batch descriptors the host planner
// (jpeg_decode.cu) fills and the three kernels (jpeg_kernels.cu) read:
//
//   jpeg_scan_*_kernel    byte-unstuffing + RSTn removal as a stream compaction over the whole batch (count per 4 KB tile, then write).
//   jpeg_entropy_kernel   one CTA per image: either one thread per restart interval, or -- files without restart markers --
//                         self-synchronising decoding: the clean scan is cut into windows, every window is decoded from a guessed
//                         state, exit states are handed forward until nothing changes, block / DC prefix sums place every thread's
//                         windows, a last pass writes the coefficients.
//   jpeg_idct_kernel      dequantise + jpeg_idct_islow, 8 lanes per block (one row / one column each), component planes out.
//   jpeg_color_kernel     fancy h2v2 / h2v1 chroma upsampling + fixed-point YCbCr -> RGB, 4 pixels per thread, into the frame buffer
//                         the letterbox and ROI kernels read (jpeg_color_rows_kernel / jpeg_color_roi_kernel: only the rows and
//                         row spans the pipeline's letterbox and warps read).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "jpeg_math.h"

namespace fdl {

constexpr int kJpegMaxBlocksPerMcu = 10;   // T.81 B.2.3
constexpr int kJpegMaxWindows = 2304;      // windows per image kept in shared memory (32 B each): two CTAs (images) fit one SM
constexpr int kJpegMinWindowBits = 1024;   // measured: 544 .. 1536 bits give the same time (3-4 hand-over rounds, each one window long)
constexpr int kJpegEntropyThreads = 512;

enum { JPEG_OK = 0, JPEG_ERR_BLOCKS = 1, JPEG_ERR_RESTARTS = 2 };

// One image of a batch.  Filled on the host from the parsed header; offsets address the batch's shared work buffers.
struct JpegImageDesc {
  long long raw_off;       // first entropy-coded byte of the scan, in the batch byte buffer
  long long clean_off;     // this image's region of the unstuffed-scan buffer (16-byte aligned, raw_len + 32 bytes)
  long long coef_off[3];   // int16 elements: component c's quantised coefficients [block rows][block cols][64], natural order
  long long plane_off[3];  // bytes: component c's sample plane [block rows * 8][block cols * 8]
  long long iv_off;        // int32 entries: restart interval starts (byte offsets in the clean scan), n_intervals of them
  long long out_off;       // bytes: where the RGB image starts in the output buffer
  int raw_len;             // bytes from raw_off to the end of the file
  int width, height, ncomp;
  int hs[3], vs[3];        // sampling factors
  int hmax, vmax, mcux, mcuy, bpm;
  int restart_interval, n_intervals;
  int blk_comp[kJpegMaxBlocksPerMcu], blk_bx[kJpegMaxBlocksPerMcu], blk_by[kJpegMaxBlocksPerMcu];
  int tab_dc[3], tab_ac[3];   // indices into the batch's JpegHuff array
  int bcols[3], brows[3];     // blocks per row / block rows of each component (MCU-padded)
  int cw[3], ch[3];           // the REAL down-sampled size ceil(image * samp / max_samp): the edges fancy upsampling replicates
  int window_bits, nwin_cap;
  int out_stride;             // bytes per output row
  int color_fast;             // 1: jpeg_color_kernel's fast path takes this image (set by the host once the output placement is known)
  uint16_t quant[3][64];      // natural order
};

// Launchers (jpeg_kernels.cu).  All buffers are the batch's; `status` receives one JPEG_* code per image.
size_t jpeg_entropy_smem_bytes(int max_windows);
int jpeg_scan_tiles(long long raw_off, int raw_len);   // 4 KB tiles of the compaction pass for one image
// compaction (two launches) + the entropy kernel; tile_info: [n][max_tiles][3] ints, scan_len: [n][2] ints of scratch
cudaError_t launch_jpeg_entropy(const JpegImageDesc* descs, int n, const JpegHuff* tabs, const uint8_t* bytes, uint8_t* clean, int16_t* coef,
                                int* iv, int* status, int* tile_info, int max_tiles, int* scan_len, int max_windows, cudaStream_t s);
// (the per-block kernel zeroes every coefficient row it has read: the buffer needs a memset only when it is new)
cudaError_t launch_jpeg_idct(const JpegImageDesc* descs, int n, int max_quads, int16_t* coef, uint8_t* planes, cudaStream_t s);
bool jpeg_idct_clears_coef();
// flags: 1 = some image takes the fast path, 2 = some image takes the generic path
cudaError_t launch_jpeg_color(const JpegImageDesc* descs, int n, int max_w, int max_h, int flags, const uint8_t* planes, uint8_t* out, cudaStream_t s);
// Sparse conversion for the pipeline (fast-path images only): the rows of a device list for every image / the row spans of the
// source quadrilaterals of `n` image_to_tensor slots (roi_stage_box + roi_row_span of glue_math.h, as roi_fill_kernel); `parents`
// != null: slot i is skipped when slot i / 2 of `parents` (its face, converted before) covers it; `rows_done` != null: one byte per
// frame row, non-zero for the rows launch_jpeg_color_rows converted.
struct I2TParams;
cudaError_t launch_jpeg_color_rows(const JpegImageDesc* descs, int n, const int* rows, int nrows, const uint8_t* planes, uint8_t* out, cudaStream_t s);
cudaError_t launch_jpeg_color_roi(const JpegImageDesc* descs, int n_images, const I2TParams* params, int n, const int* n_active,
                                  const I2TParams* parents, const uint8_t* rows_done, const uint8_t* planes, uint8_t* out, cudaStream_t s);

}  // namespace fdl
