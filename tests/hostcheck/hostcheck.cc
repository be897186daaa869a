// hostcheck.cc -- TEST INFRASTRUCTURE ONLY.  Compiles csrc/glue_math.h (the very header the CUDA kernels
// call on the device) for the host with g++ -ffp-contract=off, so that the arithmetic of the pre/post-
// processing kernels can be checked against the oracle on a CPU-only box.  Never part of libfdl_b200.so.
#include <cstring>
#include <vector>

#include "../../rs_face_detection_tflite_b200/csrc/glue_math.h"
#include "../../rs_face_detection_tflite_b200/csrc/jpeg_math.h"
#include "../../rs_face_detection_tflite_b200/csrc/jpeg_parse.h"

using namespace fdl;

extern "C" {

int hc_anchors(int model, float* out, int cap) {
  SsdOptions o;
  if (!ssd_options_for(model, &o)) return -1;
  int n = ssd_num_anchors(o);
  if (n > cap) return -2;
  for (int i = 0; i < n; ++i) ssd_anchor(o, i, &out[2 * i], &out[2 * i + 1]);
  return n;
}

// image_to_tensor on the host through the same per-pixel functions the i2t kernel uses.
int hc_image_to_tensor(const uint8_t* img, int w, int h, const fdl_rect* roi, int out_w, int out_h, int keep, double rmin, double rmax,
                       int flip, float* out, uint8_t* out_u8, double* pad4) {
  I2TParams P;
  i2t_setup(roi, w, h, out_w, out_h, keep != 0, rmin, rmax, flip != 0, 0, &P);
  if (!P.valid) return -1;
  for (int y = 0; y < out_h; ++y)
    for (int x = 0; x < out_w; ++x) {
      Px3 p = i2t_pixel(P, img_src(img, (long long)w * 3), x, y);
      size_t o = ((size_t)y * out_w + x) * 3;
      out[o] = i2t_normalise(p.r, rmin, rmax); out[o + 1] = i2t_normalise(p.g, rmin, rmax); out[o + 2] = i2t_normalise(p.b, rmin, rmax);
      out_u8[o] = (uint8_t)p.r; out_u8[o + 1] = (uint8_t)p.g; out_u8[o + 2] = (uint8_t)p.b;
    }
  for (int i = 0; i < 4; ++i) pad4[i] = P.pad[i];
  return 0;
}

// Sequential restatement of what ssd_postprocess_kernel does, built from the same scalar functions
// (decode_box, ssd_score, overlap_similarity); returns the number of output detections.
int hc_ssd_postprocess(int model, const float* reg, const float* cls, const double* pad4, fdl_detection* out, int cap, int* surv, int* n_surv) {
  SsdOptions o;
  if (!ssd_options_for(model, &o)) return -1;
  const int N = ssd_num_anchors(o);
  const float scale = (float)o.input_size;
  std::vector<int> idx; std::vector<float> score; std::vector<std::vector<float>> box;
  for (int i = 0; i < N; ++i) {
    float sc = ssd_score(cls[i]);
    if (!(sc > 0.5f)) continue;
    float ax, ay, d[16];
    ssd_anchor(o, i, &ax, &ay);
    decode_box(reg + 16 * i, ax, ay, scale, d);
    if (d[2] > d[0] && d[3] > d[1]) { idx.push_back(i); score.push_back(sc); box.emplace_back(d, d + 16); }
  }
  *n_surv = (int)idx.size();
  for (size_t i = 0; i < idx.size(); ++i) surv[i] = idx[i];
  const int n = (int)idx.size();
  std::vector<int> rem(n);
  for (int j = 0; j < n; ++j) {
    int rank = 0;
    for (int k = 0; k < n; ++k) rank += (score[k] > score[j] || (score[k] == score[j] && k < j)) ? 1 : 0;
    rem[rank] = j;
  }
  const float left = (float)pad4[0], top = (float)pad4[1];
  const float hs = (float)(1.0 - (pad4[0] + pad4[2])), vs = (float)(1.0 - (pad4[1] + pad4[3]));
  int n_out = 0;
  const double thr = (double)0.3f;
  while (!rem.empty()) {
    int t = rem[0];
    std::vector<int> cand, next;
    for (int j : rem) (overlap_similarity(box[j].data(), box[t].data()) > thr ? cand : next).push_back(j);
    float v[16];
    if (!cand.empty()) {
      for (int k = 0; k < 16; ++k) {
        float w = 0.f, total = 0.f;
        for (int c : cand) { total += score[c]; w += box[c][k] * score[c]; }
        v[k] = w / total;
      }
    } else {
      for (int k = 0; k < 16; ++k) v[k] = box[t][k];
    }
    if (n_out < cap) {
      for (int k = 0; k < 16; ++k) out[n_out].data[k] = (k & 1) ? (v[k] - top) / vs : (v[k] - left) / hs;
      out[n_out].score = score[t]; out[n_out].anchor = idx[t];
    }
    ++n_out;
    if (cand.empty()) break;
    rem.swap(next);
  }
  return n_out;
}

int hc_face_detection_to_roi(const float* data16, int w, int h, int mode, fdl_rect* out) { return face_detection_to_roi(data16, w, h, mode, out) ? 0 : -1; }
int hc_eye_roi(double ax, double ay, double bx, double by, int w, int h, fdl_rect* out) { return eye_roi(ax, ay, bx, by, w, h, out) ? 0 : -1; }
void hc_project(const float* raw, int n, int tw, int th, int iw, int ih, const double* pad4, const fdl_rect* roi, int flip, float* out) {
  ProjectParams pp;
  project_setup(tw, th, iw, ih, pad4, roi, flip != 0, &pp);
  for (int k = 0; k < n; ++k) project_point(pp, raw + 3 * k, out + 3 * k);
}

// iris refinement helpers (iris_landmark.rs:64-95, :401-433) through the same header functions the kernels call
int hc_eye_index(int eye, int* out71) {
  for (int k = 0; k < FDL_NUM_EYE_CONTOUR; ++k) out71[k] = eye_to_face_landmark_index(eye, k);
  return FDL_NUM_EYE_CONTOUR;
}
void hc_iris_metrics(const double* iris15, int w, int h, double focal, double* out2) {
  out2[0] = iris_diameter(iris15, w, h);
  out2[1] = iris_depth(iris15, focal, out2[0], w, h);
}
void hc_iris_metrics_f32(const float* iris15, int w, int h, double focal, double* out2) {
  out2[0] = iris_diameter(iris15, w, h);
  out2[1] = iris_depth(iris15, focal, out2[0], w, h);
}


// Zero-copy ROI staging (roi_fill_kernel / eye_split_kernel): every tap warp_px can read inside the frame for the face warp (and
// for an eye warp that roi_stage_covers() admits) must lie in a staged row span.  Returns the number of taps that do not (0 is
// the only acceptable answer), -1 if the face parameters are invalid, -2 if the eye is not admitted (nothing to check).
// stats[0] = staged bytes with row trimming, stats[1] = bytes of the untrimmed rectangle.
static long long uncovered_taps(const I2TParams& P, const SrcBox& b, int m) {
  long long bad = 0;
  for (int y = 0; y < P.warp_h; ++y)
    for (int x = 0; x < P.warp_w; ++x) {
      int sx, sy, ax, ay;
      warp_coords(P, x, y, &sx, &sy, &ax, &ay);
      for (int t = 0; t < 4; ++t) {
        const int tx = sx + (t & 1), ty = sy + (t >> 1);
        if (tx < 0 || tx >= P.src_w || ty < 0 || ty >= P.src_h) continue;      // border taps read nothing
        int x0, x1;
        if (!roi_row_span(b, ty, m, &x0, &x1) || tx < x0 || tx > x1) ++bad;
      }
    }
  return bad;
}
long long hc_roi_stage_check(const fdl_rect* face, const fdl_rect* eye_or_null, int w, int h, int face_size, int eye_size, int trim,
                             long long* stats) {
  I2TParams F;
  i2t_setup(face, w, h, face_size, face_size, false, 0.0, 1.0, false, 0, &F);
  if (F.valid != 1) return -1;
  int m = 0;
  const SrcBox b = roi_stage_box(F, 0, trim != 0, &m);
  if (b.x1 < b.x0) return -1;
  if (stats) {
    stats[0] = stats[1] = 0;
    for (int r = b.y0; r <= b.y1; ++r) {
      int x0, x1;
      stats[1] += (imin((3 * (b.x1 + 1) + 15) & ~15, 3 * w) - ((3 * b.x0) & ~15));
      if (roi_row_span(b, r, m, &x0, &x1)) stats[0] += (imin((3 * (x1 + 1) + 15) & ~15, 3 * w) - ((3 * x0) & ~15));
    }
  }
  if (!eye_or_null) return uncovered_taps(F, b, m);
  I2TParams E;
  i2t_setup(eye_or_null, w, h, eye_size, eye_size, true, 0.0, 1.0, false, 0, &E);
  if (E.valid != 1) return -2;
  const SrcBox e = warp_src_box(E);
  if (!roi_stage_covers(b, e, w, h)) return -2;
  return uncovered_taps(E, b, m);
}


// JPEG back half (csrc/jpeg_math.h) on the host: quantised coefficient blocks of up to three components (natural order,
// [blocks_y][blocks_x][64] per component) -> RGB, through the same per-block / per-pixel functions a device decoder will call.
// samp[c] = (h, v) sampling factors; only 1x1, 2x1 and 2x2 chroma against full-resolution luma.
int hc_jpeg_backend(int ncomp, const int16_t* const* coef, const uint16_t* const* quant, const int* samp_hv, int W, int H, uint8_t* rgb) {
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) { hmax = imax(hmax, samp_hv[2 * c]); vmax = imax(vmax, samp_hv[2 * c + 1]); }
  const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
  std::vector<std::vector<uint8_t>> planes((size_t)ncomp);
  std::vector<int> pstride((size_t)ncomp), cw((size_t)ncomp), chh((size_t)ncomp), eh((size_t)ncomp), ev((size_t)ncomp);
  for (int c = 0; c < ncomp; ++c) {
    const int h = samp_hv[2 * c], v = samp_hv[2 * c + 1];
    const int bx = mcux * h, by = mcuy * v;
    pstride[c] = bx * 8;
    planes[c].assign((size_t)bx * 8 * by * 8, 0);
    for (int j = 0; j < by; ++j)
      for (int i = 0; i < bx; ++i)
        jpeg_idct_islow_8x8(coef[c] + ((size_t)j * bx + i) * 64, quant[c], planes[c].data() + (size_t)j * 8 * pstride[c] + i * 8, pstride[c]);
    cw[c] = (W * h + hmax - 1) / hmax; chh[c] = (H * v + vmax - 1) / vmax;
    eh[c] = hmax / h; ev[c] = vmax / v;
    if (!((eh[c] == 1 && ev[c] == 1) || (eh[c] == 2 && ev[c] == 1) || (eh[c] == 2 && ev[c] == 2))) return -1;
  }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int s[3] = {0, 128, 128};
      for (int c = 0; c < ncomp; ++c) {
        if (eh[c] == 1) s[c] = planes[c][(size_t)y * pstride[c] + x];
        else if (ev[c] == 1) s[c] = jpeg_h2v1_fancy_at(planes[c].data(), pstride[c], cw[c], x, y);
        else s[c] = jpeg_h2v2_fancy_at(planes[c].data(), pstride[c], cw[c], chh[c], x, y);
      }
      uint8_t* o = rgb + ((size_t)y * W + x) * 3;
      if (ncomp == 1) { o[0] = o[1] = o[2] = (uint8_t)s[0]; }
      else jpeg_ycc_to_rgb(s[0], s[1], s[2], o);
    }
  return 0;
}


// JPEG entropy stage through csrc/jpeg_math.h: one interleaved scan, sequential over the MCUs (restart intervals honoured).
// dht[2*c] / dht[2*c+1] = (counts[16] + symbols) of component c's DC / AC table; coef[c] receives [blocks_y][blocks_x][64].
int hc_jpeg_entropy(const uint8_t* data, long long len, long long scan_offset, int ncomp, const int* samp_hv, const uint8_t* const* dht,
                    int restart_interval, int W, int H, int16_t* const* coef) {
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) { hmax = imax(hmax, samp_hv[2 * c]); vmax = imax(vmax, samp_hv[2 * c + 1]); }
  const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
  std::vector<JpegHuff> tabs((size_t)2 * ncomp);
  for (int i = 0; i < 2 * ncomp; ++i) jpeg_huff_build(dht[i], dht[i] + 16, &tabs[(size_t)i]);
  JpegBits br;
  jpeg_bits_init(&br, data + scan_offset, data + len);
  int pred[4] = {0, 0, 0, 0};
  long long count = 0;
  for (int my = 0; my < mcuy; ++my)
    for (int mx = 0; mx < mcux; ++mx) {
      if (restart_interval && count && count % restart_interval == 0) { jpeg_bits_restart(&br); pred[0] = pred[1] = pred[2] = 0; }
      ++count;
      for (int c = 0; c < ncomp; ++c) {
        const int h = samp_hv[2 * c], v = samp_hv[2 * c + 1], bxn = mcux * h;
        for (int by = 0; by < v; ++by)
          for (int bx = 0; bx < h; ++bx) {
            int16_t* blk = coef[c] + (((size_t)my * v + by) * bxn + (size_t)mx * h + bx) * 64;
            std::memset(blk, 0, 64 * sizeof(int16_t));
            jpeg_decode_block(&br, tabs[(size_t)2 * c], tabs[(size_t)2 * c + 1], &pred[c], blk);
          }
      }
    }
  return 0;
}


// The parallel schedule a device entropy stage can use when the file has restart markers: the scan is cut at the RSTn markers
// (a byte scan), and every interval is decoded on its own -- fresh bit reader at its first byte, DC predictors 0, output position
// from the interval index alone.  Intervals are visited in REVERSE order here to show that nothing flows between them.
// Returns the number of intervals, or -1 if the marker count does not match the MCU count.
int hc_jpeg_entropy_by_interval(const uint8_t* data, long long len, long long scan_offset, int ncomp, const int* samp_hv,
                                const uint8_t* const* dht, int restart_interval, int W, int H, int16_t* const* coef) {
  if (restart_interval <= 0) return -1;
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) { hmax = imax(hmax, samp_hv[2 * c]); vmax = imax(vmax, samp_hv[2 * c + 1]); }
  const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
  const long long mcus = (long long)mcux * mcuy;
  std::vector<long long> start(1, scan_offset);
  for (long long i = scan_offset; i + 1 < len; ++i) {
    if (data[i] != 0xFF) continue;
    if (data[i + 1] >= 0xD0 && data[i + 1] <= 0xD7) start.push_back(i + 2);
    else if (data[i + 1] != 0x00 && data[i + 1] != 0xFF) break;               // EOI or another segment: end of the scan
  }
  const long long intervals = (mcus + restart_interval - 1) / restart_interval;
  if ((long long)start.size() != intervals) return -1;
  std::vector<JpegHuff> tabs((size_t)2 * ncomp);
  for (int i = 0; i < 2 * ncomp; ++i) jpeg_huff_build(dht[i], dht[i] + 16, &tabs[(size_t)i]);
  for (long long iv = intervals - 1; iv >= 0; --iv) {
    JpegBits br;
    jpeg_bits_init(&br, data + start[(size_t)iv], data + len);
    int pred[4] = {0, 0, 0, 0};
    const long long m0 = iv * restart_interval, m1 = m0 + restart_interval < mcus ? m0 + restart_interval : mcus;
    for (long long m = m0; m < m1; ++m) {
      const int my = (int)(m / mcux), mx = (int)(m % mcux);
      for (int c = 0; c < ncomp; ++c) {
        const int h = samp_hv[2 * c], v = samp_hv[2 * c + 1], bxn = mcux * h;
        for (int by = 0; by < v; ++by)
          for (int bx = 0; bx < h; ++bx) {
            int16_t* blk = coef[c] + (((size_t)my * v + by) * bxn + (size_t)mx * h + bx) * 64;
            std::memset(blk, 0, 64 * sizeof(int16_t));
            jpeg_decode_block(&br, tabs[(size_t)2 * c], tabs[(size_t)2 * c + 1], &pred[c], blk);
          }
      }
    }
  }
  return (int)intervals;
}


// Whole decode on the host through csrc/jpeg_parse.h + csrc/jpeg_math.h: bytes -> RGB.  Two calls: rgb == nullptr returns the
// size; returns 0, or -1 with the parser's message in err.
int hc_jpeg_decode(const uint8_t* data, long long len, uint8_t* rgb, int* w, int* h, char* err, int errcap) {
  JpegHeader hd;
  std::string msg;
  if (!jpeg_parse_header(data, (size_t)len, &hd, &msg)) {
    if (err && errcap > 0) { std::strncpy(err, msg.c_str(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return -1;
  }
  *w = hd.width; *h = hd.height;
  if (!rgb) return 0;
  int samp[6];
  const uint8_t* dht[6];
  const uint16_t* quant[3];
  std::vector<std::vector<int16_t>> coef((size_t)hd.ncomp);
  int16_t* cp[3];
  for (int c = 0; c < hd.ncomp; ++c) {
    samp[2 * c] = hd.comp[c].h; samp[2 * c + 1] = hd.comp[c].v;
    dht[2 * c] = hd.dht[0][hd.comp[c].td]; dht[2 * c + 1] = hd.dht[1][hd.comp[c].ta];
    quant[c] = hd.quant[hd.comp[c].tq];
    coef[(size_t)c].assign((size_t)hd.mcus_x * hd.comp[c].h * hd.mcus_y * hd.comp[c].v * 64, 0);
    cp[c] = coef[(size_t)c].data();
  }
  hc_jpeg_entropy(data, len, (long long)hd.scan_offset, hd.ncomp, samp, dht, hd.restart_interval, hd.width, hd.height, cp);
  const int16_t* ccp[3] = {cp[0], hd.ncomp > 1 ? cp[1] : nullptr, hd.ncomp > 2 ? cp[2] : nullptr};
  return hc_jpeg_backend(hd.ncomp, ccp, quant, samp, hd.width, hd.height, rgb);
}


// The self-synchronising schedule (jpeg_math.h) on the host, for scans WITHOUT restart markers: unstuff, cut into windows of
// `window_bits`, decode every window from a guessed entry state, hand exit states forward until a fixed point (every round is
// one parallel step on a device), then one output pass with block offsets from a prefix sum and DC values from a per-component
// prefix sum of the differences.  Returns the number of rounds the fixed point took (>= 1), or -1.
int hc_jpeg_entropy_selfsync(const uint8_t* data, long long len, long long scan_offset, int ncomp, const int* samp_hv, const uint8_t* const* dht,
                             int W, int H, int window_bits, int16_t* const* coef, int* windows_out, int* redecoded_out) {
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; ++c) { hmax = imax(hmax, samp_hv[2 * c]); vmax = imax(vmax, samp_hv[2 * c + 1]); }
  const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
  uint8_t comp_of_block[16];
  int bpm = 0;
  for (int c = 0; c < ncomp; ++c) for (int i = 0; i < samp_hv[2 * c] * samp_hv[2 * c + 1]; ++i) { if (bpm >= 16) return -1; comp_of_block[bpm++] = (uint8_t)c; }
  const long long total_blocks = (long long)mcux * mcuy * bpm;
  std::vector<JpegHuff> tabs((size_t)2 * ncomp);
  for (int i = 0; i < 2 * ncomp; ++i) jpeg_huff_build(dht[i], dht[i] + 16, &tabs[(size_t)i]);
  // unstuff (a stream compaction on a device)
  std::vector<uint8_t> clean;
  for (long long i = scan_offset; i < len; ++i) {
    if (data[i] == 0xFF) { if (i + 1 < len && data[i + 1] == 0x00) { clean.push_back(0xFF); ++i; continue; } break; }
    clean.push_back(data[i]);
  }
  const long long nbits = (long long)clean.size() * 8;
  const int nwin = (int)((nbits + window_bits - 1) / window_bits);
  if (nwin < 1) return -1;
  std::vector<JpegSyncState> entry((size_t)nwin), exit_((size_t)nwin);
  std::vector<long long> blocks((size_t)nwin, 0);
  auto run = [&](int i, bool write, long long block0) {
    JpegSyncState st = entry[(size_t)i];
    const long long end = (long long)(i + 1) * window_bits < nbits ? (long long)(i + 1) * window_bits : nbits;
    long long nb = 0;
    while (st.pos < end) {
      int zz, value; bool done;
      const int b = st.b;
      jpeg_sync_step(clean.data(), nbits, tabs.data(), comp_of_block, bpm, &st, &zz, &value, &done);
      if (write && zz >= 0) {
        const long long g = block0 + nb;
        if (g < total_blocks) {
          const long long m = g / bpm;
          const int c = comp_of_block[b], h = samp_hv[2 * c], v = samp_hv[2 * c + 1];
          int first = 0; while (comp_of_block[first] != c) ++first;
          const int j = b - first, by = j / h, bx = j % h, my = (int)(m / mcux), mx = (int)(m % mcux);
          static const uint8_t zzt[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
          coef[c][((((size_t)my * v + by) * ((size_t)mcux * h)) + (size_t)mx * h + bx) * 64 + zzt[zz]] = (int16_t)value;   // DC: the difference for now
        }
      }
      if (done) ++nb;
    }
    exit_[(size_t)i] = st; blocks[(size_t)i] = nb;
  };
  for (int i = 0; i < nwin; ++i) { entry[(size_t)i].pos = (long long)i * window_bits; entry[(size_t)i].b = 0; entry[(size_t)i].k = 0; run(i, false, 0); }
  int rounds = 1, redecoded = 0;
  for (;; ++rounds) {
    std::vector<JpegSyncState> next(entry);
    for (int i = 1; i < nwin; ++i) next[(size_t)i] = exit_[(size_t)i - 1];        // all from the previous round: one parallel step
    bool changed = false;
    for (int i = 1; i < nwin; ++i)
      if (!(next[(size_t)i] == entry[(size_t)i])) { entry[(size_t)i] = next[(size_t)i]; run(i, false, 0); changed = true; ++redecoded; }
    if (!changed) break;
    if (rounds > nwin + 2) return -1;
  }
  // output pass
  for (int c = 0; c < ncomp; ++c) std::memset(coef[c], 0, (size_t)mcux * samp_hv[2 * c] * mcuy * samp_hv[2 * c + 1] * 64 * sizeof(int16_t));
  long long before = 0;
  for (int i = 0; i < nwin; ++i) { const long long nb = blocks[(size_t)i]; run(i, true, before); before += nb; }
  if (before < total_blocks) return -1;
  // DC differences -> values: per-component running sum in scan order
  int pred[4] = {0, 0, 0, 0};
  for (int my = 0; my < mcuy; ++my)
    for (int mx = 0; mx < mcux; ++mx)
      for (int c = 0; c < ncomp; ++c) {
        const int h = samp_hv[2 * c], v = samp_hv[2 * c + 1];
        for (int by = 0; by < v; ++by)
          for (int bx = 0; bx < h; ++bx) {
            int16_t* blk = coef[c] + ((((size_t)my * v + by) * ((size_t)mcux * h)) + (size_t)mx * h + bx) * 64;
            pred[c] += blk[0];
            blk[0] = (int16_t)pred[c];
          }
      }
  if (windows_out) *windows_out = nwin;
  if (redecoded_out) *redecoded_out = redecoded;
  return rounds;
}

}
