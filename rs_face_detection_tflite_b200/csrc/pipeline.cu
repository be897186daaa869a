// pipeline.cu -- batched detect -> face ROI -> landmark -> eye ROIs -> iris(L,R) on one B200.
//
// The call sequence is the reference's canonical one (lib.rs:20-40): FaceDetection::infer ->
// face_detection_to_roi(faces[i]) -> FaceLandmark::infer -> iris_roi_from_face_landmarks ->
// IrisLandmark::infer (right, left).  Here it runs for a whole batch of frames without leaving the
// device: the data-dependent fan-out (0..max_faces faces per frame, two eyes per face) is handled
// with device-side slot lists and counters, never with a host round trip.
//
// Concurrency: `kDepth` lanes, each with its own stream, frame buffer, network-input buffers, ROI / slot
// scratch and result buffers.  A lane's stream runs copy-in (or, in zero-copy mode, kernels that read the
// pinned host frames in place over PCIe), all kernels and the copy-out of one batch in order; two lanes in
// flight overlap one batch's PCIe traffic with the other's compute.  The three networks (weights +
// activation arenas) are shared: a per-network event serialises their use across lanes.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "device_util.h"
#include "fdl_status.h"
#include "glue_math.h"
#include "jpeg_decode.h"
#include "net.h"
#include "prepost_kernels.cuh"

using namespace fdl;

namespace {
constexpr int kDepth = 4;
constexpr int kStages = 10;

// Zero-copy host mode: the detector's letterbox (image_to_tensor with roi = None, transform.rs:239-280) interpolates
// between only a few source rows (2 of every 7.5 for 1080p -> 256).  Those rows form a periodic pattern, so the COPY
// ENGINE can gather exactly them with a handful of strided 2-D copies (no SM involved, full PCIe rate, overlaps any
// kernel); the letterbox kernel then reads the compact copy from HBM.  The ROI warps keep reading the pinned frames
// in place.
struct RowFamily { int src_row0, rows, dst_row0; };
struct RowGather {
  bool ok = false;
  int rows_per_frame = 0;      // compact rows per frame
  int period_src_rows = 0;     // source rows per period (the pattern repeats every so many rows)
  int periods_per_frame = 0;
  int rows_per_period = 0;     // compact rows per period
  std::vector<RowFamily> fam;  // one strided copy per family
  DevBuf<int> row_pos;         // [H] source row -> compact row (-1: not gathered)
};

// The frame rows the detector's letterbox reads (the plain-letterbox path of i2t_rows_kernel: two source rows per output row), or
// false when the slot transform is not the plain letterbox.
bool letterbox_rows(int W, int H, int S, std::vector<int>* rows) {
  I2TParams P;
  i2t_setup(nullptr, W, H, S, S, true, -1.0, 1.0, false, 0, &P);
  const int bw = P.warp_w + 2 * P.pad_h, bh = P.warp_h + 2 * P.pad_v;
  const bool simple = P.valid && P.has_r2 && !P.flip && P.warp_w == P.src_w && P.warp_h == P.src_h && (!P.has_r1 || (bw == P.r1_w && bh == P.r1_h)) &&
                      !(P.r1_w == S && P.r1_h == S);
  if (!simple) return false;
  // the kernel's own test of the identity warp (i2t_rows_kernel): a slot that fails it is evaluated pixel by pixel anywhere in the frame
  const double e = 1e-9;
  if (!(fabs(P.Mi[0] - 1.0) < e && fabs(P.Mi[4] - 1.0) < e && fabs(P.Mi[8] - 1.0) < e && fabs(P.Mi[1]) < e && fabs(P.Mi[2]) < e * P.src_w &&
        fabs(P.Mi[3]) < e && fabs(P.Mi[5]) < e * P.src_h && fabs(P.Mi[6]) < e && fabs(P.Mi[7]) < e))
    return false;
  const int pv = P.has_r1 ? P.pad_v : 0;
  std::vector<char> need((size_t)H, 0);
  for (int oy = 0; oy < S; ++oy) {
    int y0, y1, b0, b1;
    resize_coeff(oy, S, P.r1_h, false, &y0, &y1, &b0, &b1);
    const int sy0 = y0 - pv, sy1 = y1 - pv;
    if (sy0 >= 0 && sy0 < H) need[(size_t)sy0] = 1;
    if (sy1 >= 0 && sy1 < H) need[(size_t)sy1] = 1;
  }
  rows->clear();
  for (int r = 0; r < H; ++r) if (need[(size_t)r]) rows->push_back(r);
  return !rows->empty();
}

bool plan_row_gather(int W, int H, int S, RowGather* g, std::vector<int>* row_pos_host) {
  I2TParams P;
  i2t_setup(nullptr, W, H, S, S, true, -1.0, 1.0, false, 0, &P);
  const int bw = P.warp_w + 2 * P.pad_h, bh = P.warp_h + 2 * P.pad_v;
  const bool simple = P.valid && P.has_r2 && !P.flip && P.warp_w == P.src_w && P.warp_h == P.src_h && (!P.has_r1 || (bw == P.r1_w && bh == P.r1_h)) &&
                      !(P.r1_w == S && P.r1_h == S);
  if (!simple) return false;
  const int pv = P.has_r1 ? P.pad_v : 0;
  std::vector<char> need((size_t)H, 0);
  for (int oy = 0; oy < S; ++oy) {
    int y0, y1, b0, b1;
    resize_coeff(oy, S, P.r1_h, false, &y0, &y1, &b0, &b1);
    const int sy0 = y0 - pv, sy1 = y1 - pv;
    if (sy0 >= 0 && sy0 < H) need[(size_t)sy0] = 1;
    if (sy1 >= 0 && sy1 < H) need[(size_t)sy1] = 1;
  }
  std::vector<int> start, len;
  for (int r = 0; r < H;) {
    if (!need[(size_t)r]) { ++r; continue; }
    int e = r;
    while (e < H && need[(size_t)e]) ++e;
    start.push_back(r); len.push_back(e - r);
    r = e;
  }
  const int nruns = (int)start.size();
  int total = 0;
  for (int l : len) total += l;
  if (nruns == 0 || total * 10 > H * 6) return false;     // most of the frame is needed: nothing to gain
  for (int p = 1; p <= 8 && p <= nruns; ++p) {
    if (nruns % p) continue;
    bool good = true;
    const int D = nruns > p ? start[(size_t)p] - start[0] : H;
    for (int i = 0; i + p < nruns && good; ++i) good = start[(size_t)(i + p)] - start[(size_t)i] == D && len[(size_t)(i + p)] == len[(size_t)i];
    if (!good || (long long)D * (nruns / p) != H) continue;   // the pattern must continue seamlessly into the next frame
    g->fam.clear();
    int prefix = 0;
    for (int k = 0; k < p; ++k) { g->fam.push_back({start[(size_t)k], len[(size_t)k], prefix}); prefix += len[(size_t)k]; }
    g->rows_per_period = prefix; g->period_src_rows = D; g->periods_per_frame = nruns / p; g->rows_per_frame = prefix * (nruns / p);
    row_pos_host->assign((size_t)H, -1);
    for (int i = 0; i < nruns; ++i)
      for (int j = 0; j < len[(size_t)i]; ++j) (*row_pos_host)[(size_t)(start[(size_t)i] + j)] = (i / p) * prefix + g->fam[(size_t)(i % p)].dst_row0 + j;
    g->ok = true;
    return true;
  }
  return false;
}

struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> frames;
  DevBuf<uint8_t> rows;                        // copy-engine gather of the letterbox source rows (zero-copy host mode)
  DevBuf<float> det_in, lmk_in, iris_in;       // image_to_tensor outputs (network inputs), private to the lane
  DevBuf<I2TParams> det_params, face_params, eye_params, eye_params_dev, eye_params_host;
  DevBuf<SrcBox> face_boxes;                   // zero-copy host mode: the frame rectangle staged on the device per face slot
  DevBuf<fdl_rect> face_rois, eye_rois;
  DevBuf<int> slot_frame, slot_face, face_valid, eye_frame, eye_valid, counters;
  DevBuf<fdl_frame_result> d_frames;
  DevBuf<fdl_face_result> d_faces;
  PinBuf<fdl_frame_result> h_frames;
  PinBuf<fdl_face_result> h_faces;
  cudaEvent_t ev_h2d_start = nullptr, ev_stage[kStages] = {}, ev_done = nullptr;
  JpegDecoder jpeg;                            // fdl_pipeline_submit_jpeg: the lane's device decoder (frames land in `frames`)
  bool jpeg_pending = false;
  int n = 0;
  int ticket = -1;
  bool busy = false;
};
}  // namespace

struct fdl_pipeline {
  fdl_pipeline_config cfg{};
  std::string model_dir;
  Net* det = nullptr;
  Net* lmk = nullptr;
  Net* iris = nullptr;
  cudaEvent_t guard[3] = {};     // last use of det / lmk / iris by any lane
  SsdOptions opt{};
  int S = 0, N = 0, LS = 0, IS = 0;
  DevBuf<float> anchors;
  RowGather gather;
  DevBuf<uint8_t> jpeg_rows_done;   // the same rows as a byte mask over the frame rows
  DevBuf<int> jpeg_rows;       // frame rows the letterbox reads (sparse colour conversion of JPEG batches); empty: convert whole frames
  int n_jpeg_rows = 0;
  Lane lanes[kDepth];
  // One thread may submit while another collects (fdl_pool's worker and the application thread): the lane table is guarded, and
  // the guard is NOT held while collect waits for the GPU.
  std::mutex mu;
  int next_ticket = 0;
  float last_device_ms = 0.f;
  float stage_ms[kStages] = {};
};

static void pipeline_free(fdl_pipeline* p) {
  if (!p) return;
  DeviceGuard _device_guard;
  cudaSetDevice(p->cfg.device);
  cudaDeviceSynchronize();
  for (auto& l : p->lanes) {
    if (l.ev_h2d_start) cudaEventDestroy(l.ev_h2d_start);
    if (l.ev_done) cudaEventDestroy(l.ev_done);
    for (auto& e : l.ev_stage) if (e) cudaEventDestroy(e);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  for (auto& g : p->guard) if (g) cudaEventDestroy(g);
  delete p->det; delete p->lmk; delete p->iris;
  delete p;
}

extern "C" {

int fdl_pipeline_create(const fdl_pipeline_config* cfg, fdl_pipeline** out) try {
  DeviceGuard _device_guard;
  if (!cfg || !out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->max_batch <= 0 || cfg->max_faces <= 0 || cfg->max_faces > FDL_MAX_DETECTIONS || cfg->frame_width <= 0 || cfg->frame_height <= 0)
    return set_error(FDL_ERR_INVALID, "bad pipeline configuration");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(FDL_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(FDL_ERR_INVALID, "device index out of range");
  FDL_CUDA_TRY(cudaSetDevice(cfg->device));
  fdl_pipeline* p = new fdl_pipeline();
  p->cfg = *cfg;
  p->model_dir = cfg->model_dir ? cfg->model_dir : "./models";
  p->cfg.model_dir = nullptr;
  auto bail = [&](int c, const std::string& m) { pipeline_free(p); return set_error(c, m); };
  const char* file = nullptr;
  switch (cfg->detector_model) {
    case FDL_MODEL_FRONT_CAMERA: file = "face_detection_front.tflite"; break;
    case FDL_MODEL_BACK_CAMERA: file = "face_detection_back.tflite"; break;
    case FDL_MODEL_SHORT: file = "face_detection_short_range.tflite"; break;
    case FDL_MODEL_FULL: file = "face_detection_full_range.tflite"; break;
    case FDL_MODEL_FULL_SPARSE: file = "face_detection_full_range_sparse.tflite"; break;   // face_detection.rs:180-183
    default: return bail(FDL_ERR_MODEL, "unsupported model type");
  }
  ssd_options_for(cfg->detector_model, &p->opt);
  std::string err; int code = FDL_ERR_INTERNAL;
  p->det = Net::create(p->model_dir + "/" + file, cfg->device, &err, &code);
  if (!p->det) return bail(code, err);
  p->S = p->det->plan().input.H; p->N = ssd_num_anchors(p->opt);
  if (p->det->num_outputs() != 2 || p->det->out_elems(0) != (int64_t)p->N * 16 || p->det->out_elems(1) != p->N || p->S != p->opt.input_size)
    return bail(FDL_ERR_MODEL, "incompatible detector model");
  if (cfg->run_landmarks) {
    p->lmk = Net::create(p->model_dir + "/face_landmark.tflite", cfg->device, &err, &code);
    if (!p->lmk) return bail(code, err);
    p->LS = p->lmk->plan().input.H;
    if (p->lmk->num_outputs() != 2 || p->lmk->out_elems(0) < 3 * FDL_NUM_FACE_LANDMARKS) return bail(FDL_ERR_MODEL, "incompatible landmark model");
    if (cfg->run_iris) {
      p->iris = Net::create(p->model_dir + "/iris_landmark.tflite", cfg->device, &err, &code);
      if (!p->iris) return bail(code, err);
      p->IS = p->iris->plan().input.H;
      if (p->iris->num_outputs() != 2 || p->iris->out_elems(0) != 3 * FDL_NUM_EYE_CONTOUR || p->iris->out_elems(1) != 3 * FDL_NUM_IRIS)
        return bail(FDL_ERR_MODEL, "incompatible iris model");
    }
  }
  const int B = cfg->max_batch, F = B * cfg->max_faces, E = 2 * F;
  cudaError_t e = p->anchors.reserve((size_t)p->N * 2);
  for (auto& g : p->guard) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g, cudaEventDisableTiming);
  const size_t frame_bytes = (size_t)cfg->frame_width * 3 * cfg->frame_height;
  for (auto& l : p->lanes) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = l.frames.reserve(frame_bytes * B);
    if (e == cudaSuccess) e = l.det_in.reserve((size_t)B * p->S * p->S * 3);
    if (e == cudaSuccess && p->lmk) e = l.lmk_in.reserve((size_t)F * p->LS * p->LS * 3);
    if (e == cudaSuccess && p->iris) e = l.iris_in.reserve((size_t)E * p->IS * p->IS * 3);
    if (e == cudaSuccess) e = l.det_params.reserve(B);
    if (e == cudaSuccess) e = l.face_params.reserve(F);
    if (e == cudaSuccess) e = l.eye_params.reserve(E);
    if (e == cudaSuccess && cfg->zero_copy_host) e = l.eye_params_dev.reserve(E);
    if (e == cudaSuccess && cfg->zero_copy_host) e = l.eye_params_host.reserve(E);
    if (e == cudaSuccess && cfg->zero_copy_host) e = l.face_boxes.reserve(F);
    if (e == cudaSuccess) e = l.face_rois.reserve(F);
    if (e == cudaSuccess) e = l.eye_rois.reserve(E);
    if (e == cudaSuccess) e = l.slot_frame.reserve(F);
    if (e == cudaSuccess) e = l.slot_face.reserve(F);
    if (e == cudaSuccess) e = l.face_valid.reserve(F);
    if (e == cudaSuccess) e = l.eye_frame.reserve(E);
    if (e == cudaSuccess) e = l.eye_valid.reserve(E);
    if (e == cudaSuccess) e = l.counters.reserve(4);
    if (e == cudaSuccess) e = cudaMemsetAsync(l.counters.p, 0, 4 * sizeof(int), l.stream);
    if (e == cudaSuccess) e = l.d_frames.reserve(B);
    if (e == cudaSuccess) e = l.d_faces.reserve(F);
    if (e == cudaSuccess) e = l.h_frames.reserve(B);
    if (e == cudaSuccess) e = l.h_faces.reserve(F);
    if (e == cudaSuccess) e = cudaEventCreate(&l.ev_h2d_start);
    if (e == cudaSuccess) e = cudaEventCreate(&l.ev_done);
    for (auto& ev : l.ev_stage) if (e == cudaSuccess) e = cudaEventCreate(&ev);
  }
  if (e == cudaSuccess && cfg->zero_copy_host) {
    std::vector<int> rp;
    if (plan_row_gather(cfg->frame_width, cfg->frame_height, p->S, &p->gather, &rp)) {
      e = p->gather.row_pos.reserve(rp.size());
      if (e == cudaSuccess) e = cudaMemcpy(p->gather.row_pos.p, rp.data(), rp.size() * sizeof(int), cudaMemcpyHostToDevice);
      for (auto& l : p->lanes)
        if (e == cudaSuccess) e = l.rows.reserve((size_t)B * p->gather.rows_per_frame * cfg->frame_width * 3);
    }
  }
  {
    // JPEG batches: nobody but the letterbox and the ROI warps reads the decoded frames, so only what they read is colour-converted
    static const int lazy_env = getenv("FDL_JPEG_SPARSE") ? atoi(getenv("FDL_JPEG_SPARSE")) : 1;
    std::vector<int> rows;
    if (e == cudaSuccess && lazy_env && letterbox_rows(cfg->frame_width, cfg->frame_height, p->S, &rows) &&
        rows.size() * 10 <= (size_t)cfg->frame_height * 6) {     // (small frames: most rows are read, nothing to gain)
      e = p->jpeg_rows.reserve(rows.size());
      if (e == cudaSuccess) e = cudaMemcpy(p->jpeg_rows.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice);
      std::vector<uint8_t> mask((size_t)cfg->frame_height, 0);
      for (int r : rows) mask[(size_t)r] = 1;
      if (e == cudaSuccess) e = p->jpeg_rows_done.reserve(mask.size());
      if (e == cudaSuccess) e = cudaMemcpy(p->jpeg_rows_done.p, mask.data(), mask.size(), cudaMemcpyHostToDevice);
      p->n_jpeg_rows = e == cudaSuccess ? (int)rows.size() : 0;
    }
  }
  if (e == cudaSuccess) e = launch_anchors(p->opt, p->anchors.p, p->N, p->lanes[0].stream);
  if (e != cudaSuccess) return bail(FDL_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
  if (!p->det->reserve(B, &err) || (p->lmk && !p->lmk->reserve(F, &err)) || (p->iris && !p->iris->reserve(E, &err)))
    return bail(FDL_ERR_CUDA, err);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return bail(FDL_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
  *out = p;
  return FDL_OK;
} FDL_ABI_CATCH

void fdl_pipeline_destroy(fdl_pipeline* p) { pipeline_free(p); }
int fdl_pipeline_depth(const fdl_pipeline*) { return kDepth; }

}  // extern "C"

// Everything after the frames are where the kernels can read them (`fptr`: the lane's frame buffer, the caller's device frames, or --
// zero-copy host mode -- the caller's pinned frames, `host_base` their host address for the copy engine's row gather).
static int pipeline_enqueue(fdl_pipeline* p, Lane* lane, int n, const uint8_t* fptr, bool used_host, const uint8_t* host_base, int* ticket,
                            JpegDecoder* sparse_jpeg = nullptr) {
  const int W = p->cfg.frame_width, H = p->cfg.frame_height, MF = p->cfg.max_faces;
  cudaStream_t cs = lane->stream;
  const long long row = (long long)W * 3, fstride = row * H;
  static const int zc_env = getenv("FDL_ZC_CTAS") ? atoi(getenv("FDL_ZC_CTAS")) : 148;
  const int zc_ctas = used_host ? zc_env : 0;   // persistent CTAs for kernels that read host memory over PCIe
  const bool gathered = used_host && p->gather.ok;
  if (gathered) {
    // the copy engine gathers the source rows of the letterbox: one strided 2-D copy per row family for the whole batch
    const RowGather& g = p->gather;
    for (const RowFamily& f : g.fam)
      FDL_CUDA_TRY(cudaMemcpy2DAsync(lane->rows.p + (size_t)f.dst_row0 * row, (size_t)g.rows_per_period * row, host_base + (size_t)f.src_row0 * row,
                                     (size_t)g.period_src_rows * row, (size_t)f.rows * row, (size_t)n * g.periods_per_frame, cudaMemcpyHostToDevice, cs));
  }
  const int F = n * MF, E = 2 * F;
  int* n_faces = lane->counters.p;
  int* n_eyes = lane->counters.p + 1;
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[0], cs));
  FDL_CUDA_TRY(cudaMemsetAsync(lane->d_frames.p, 0, (size_t)n * sizeof(fdl_frame_result), cs));
  FDL_CUDA_TRY(cudaMemsetAsync(lane->d_faces.p, 0, (size_t)F * sizeof(fdl_face_result), cs));
  // FaceDetection::infer: image_to_tensor(keep_aspect, (-1,1)) -> net -> SSD post-processing
  FDL_CUDA_TRY(launch_i2t_setup(nullptr, nullptr, nullptr, n, W, H, p->S, p->S, 1, -1.0, 1.0, 0, lane->det_params.p, nullptr, cs));
  FDL_CUDA_TRY(launch_i2t(fptr, fstride, row, lane->det_params.p, n, p->S, p->S, lane->det_in.p, (long long)p->S * p->S * 3, nullptr, nullptr, cs, 1, W,
                          gathered ? 0 : zc_ctas, gathered ? lane->rows.p : nullptr, gathered ? p->gather.row_pos.p : nullptr,
                          gathered ? (long long)p->gather.rows_per_frame * row : 0));
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[1], cs));
  FDL_CUDA_TRY(cudaStreamWaitEvent(cs, p->guard[0], 0));
  FDL_CUDA_TRY(p->det->forward(n, cs, nullptr, lane->det_in.p));
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[2], cs));
  {
    TView reg = p->det->output_view(0, n), cls = p->det->output_view(1, n);
    SsdPostArgs a;
    a.reg = reg.p; a.reg_bstride = reg.bstride; a.cls = cls.p; a.cls_bstride = cls.bstride;
    a.anchors = p->anchors.p; a.N = p->N; a.B = n; a.scale = (float)p->S; a.params = lane->det_params.p;
    a.det_base = reinterpret_cast<char*>(lane->d_frames.p) + offsetof(fdl_frame_result, detections);
    a.det_stride = sizeof(fdl_frame_result);
    a.ndet_base = reinterpret_cast<char*>(lane->d_frames.p) + offsetof(fdl_frame_result, n_detections);
    a.ndet_stride = sizeof(fdl_frame_result);
    a.max_out = FDL_MAX_DETECTIONS;
    static_assert(sizeof(fdl_frame_result) % sizeof(int) == 0 && offsetof(fdl_frame_result, n_total_detections) % sizeof(int) == 0, "int-strided counters");
    a.n_total = reinterpret_cast<int*>(reinterpret_cast<char*>(lane->d_frames.p) + offsetof(fdl_frame_result, n_total_detections));
    a.n_total_stride = sizeof(fdl_frame_result) / sizeof(int);
    FDL_CUDA_TRY(launch_ssd_postprocess(a, cs));
  }
  FDL_CUDA_TRY(cudaEventRecord(p->guard[0], cs));
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[3], cs));
  if (p->lmk) {
    // face_detection_to_roi + FaceLandmark::infer for the first max_faces detections of every frame
    FDL_CUDA_TRY(launch_face_select(lane->d_frames.p, n, MF, lane->slot_frame.p, lane->slot_face.p, n_faces, n_eyes, cs));
    FDL_CUDA_TRY(launch_face_roi(lane->d_frames.p, lane->slot_frame.p, lane->slot_face.p, F, MF, W, H, lane->face_rois.p, lane->face_valid.p,
                                 lane->d_faces.p, n_faces, cs));
    FDL_CUDA_TRY(launch_i2t_setup(lane->face_rois.p, lane->slot_frame.p, lane->face_valid.p, F, W, H, p->LS, p->LS, 0, 0.0, 1.0, 0,
                                  lane->face_params.p, n_faces, cs));
    // zero-copy host frames: stage the faces' source rectangles on the device once; the face warp and (normally) both eye warps
    // then read the device copy instead of fetching their taps over PCIe
    // sparsely converted JPEG frames: the source regions of the face warps are converted now that they are known
    if (sparse_jpeg) FDL_CUDA_TRY(sparse_jpeg->color_roi(lane->face_params.p, F, n_faces, nullptr, cs));
    static const int crop_env = getenv("FDL_ZC_CROP") ? atoi(getenv("FDL_ZC_CROP")) : 1;
    const bool crop = used_host && crop_env && lane->face_boxes.p && lane->frames.cap >= (size_t)fstride * n && (row & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(fptr) & 15) == 0;     // 16-byte row copies
    if (crop) {
      // rows of the rectangle that the letterbox gather already brought to the device (2 of every 7.5 at 1080p) are taken from there
      static const int reuse_env = getenv("FDL_ZC_REUSE") ? atoi(getenv("FDL_ZC_REUSE")) : 1;
      const bool reuse = gathered && reuse_env;
      FDL_CUDA_TRY(launch_roi_fill(fptr, lane->frames.p, fstride, row, lane->face_params.p, F, n_faces, lane->face_boxes.p, H, zc_ctas * 4, cs,
                                   reuse ? lane->rows.p : nullptr, reuse ? p->gather.row_pos.p : nullptr,
                                   reuse ? (long long)p->gather.rows_per_frame * row : 0));
      FDL_CUDA_TRY(launch_i2t(lane->frames.p, fstride, row, lane->face_params.p, F, p->LS, p->LS, lane->lmk_in.p, (long long)p->LS * p->LS * 3, nullptr, n_faces, cs, 0, 0, 0));
    } else {
      FDL_CUDA_TRY(launch_i2t(fptr, fstride, row, lane->face_params.p, F, p->LS, p->LS, lane->lmk_in.p, (long long)p->LS * p->LS * 3, nullptr, n_faces, cs, 0, 0, zc_ctas));
    }
    FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[4], cs));
    FDL_CUDA_TRY(cudaStreamWaitEvent(cs, p->guard[1], 0));
    FDL_CUDA_TRY(p->lmk->forward(F, cs, n_faces, lane->lmk_in.p));
    FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[5], cs));
    TView raw = p->lmk->output_view(0, F), flag = p->lmk->output_view(1, F);
    // the reference reads the LAST element of the flag tensor (face_landmark.rs:292-293)
    FDL_CUDA_TRY(launch_landmark_post(raw.p, raw.bstride, flag.p + p->lmk->out_elems(1) - 1, flag.bstride, lane->face_params.p, lane->face_rois.p,
                                      lane->slot_frame.p, lane->slot_face.p, F, MF, p->LS, p->LS, lane->d_faces.p, lane->eye_rois.p,
                                      lane->eye_frame.p, lane->eye_valid.p, n_faces, cs, (p->iris && p->cfg.refine_landmarks) ? 1 : 0));
    FDL_CUDA_TRY(cudaEventRecord(p->guard[1], cs));
    if (p->iris) {
      // IrisLandmark::infer for both eyes: image_to_tensor(keep_aspect, (0,1), flip = right eye)
      FDL_CUDA_TRY(launch_i2t_setup(lane->eye_rois.p, lane->eye_frame.p, lane->eye_valid.p, E, W, H, p->IS, p->IS, 1, 0.0, 1.0, 2,
                                    lane->eye_params.p, n_eyes, cs));
      if (sparse_jpeg) FDL_CUDA_TRY(sparse_jpeg->color_roi(lane->eye_params.p, E, n_eyes, lane->face_params.p, cs));     // (an eye ROI may leave its face's)
      if (crop) {
        FDL_CUDA_TRY(launch_eye_split(lane->eye_params.p, lane->face_boxes.p, E, n_eyes, lane->eye_params_dev.p, lane->eye_params_host.p, cs));
        FDL_CUDA_TRY(launch_i2t(lane->frames.p, fstride, row, lane->eye_params_dev.p, E, p->IS, p->IS, lane->iris_in.p, (long long)p->IS * p->IS * 3, nullptr, n_eyes, cs, 0, 0, 0));
        FDL_CUDA_TRY(launch_i2t(fptr, fstride, row, lane->eye_params_host.p, E, p->IS, p->IS, lane->iris_in.p, (long long)p->IS * p->IS * 3, nullptr, n_eyes, cs, 0, 0, zc_ctas));
      } else {
        FDL_CUDA_TRY(launch_i2t(fptr, fstride, row, lane->eye_params.p, E, p->IS, p->IS, lane->iris_in.p, (long long)p->IS * p->IS * 3, nullptr, n_eyes, cs, 0, 0, zc_ctas));
      }
      FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[6], cs));
      FDL_CUDA_TRY(cudaStreamWaitEvent(cs, p->guard[2], 0));
      FDL_CUDA_TRY(p->iris->forward(E, cs, n_eyes, lane->iris_in.p));
      FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[7], cs));
      TView ec = p->iris->output_view(0, E), ir = p->iris->output_view(1, E);
      FDL_CUDA_TRY(launch_iris_post(ec.p, ec.bstride, ir.p, ir.bstride, lane->eye_params.p, lane->eye_rois.p, lane->eye_valid.p, lane->slot_frame.p,
                                    lane->slot_face.p, E, MF, p->IS, p->IS, lane->d_faces.p, n_eyes, cs, p->cfg.refine_landmarks ? 1 : 0,
                                    p->cfg.focal_length_mm));
      FDL_CUDA_TRY(cudaEventRecord(p->guard[2], cs));
    } else {
      FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[6], cs));
      FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[7], cs));
    }
  } else {
    for (int i = 4; i <= 7; ++i) FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[i], cs));
  }
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_stage[8], cs));

  // ---- copy-out
  FDL_CUDA_TRY(cudaMemcpyAsync(lane->h_frames.p, lane->d_frames.p, (size_t)n * sizeof(fdl_frame_result), cudaMemcpyDeviceToHost, cs));
  if (p->lmk)
    FDL_CUDA_TRY(cudaMemcpyAsync(lane->h_faces.p, lane->d_faces.p, (size_t)F * sizeof(fdl_face_result), cudaMemcpyDeviceToHost, cs));
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_done, cs));
  lane->n = n;
  lane->busy = true;
  lane->ticket = p->next_ticket++;
  *ticket = lane->ticket;
  return FDL_OK;
}

extern "C" {

int fdl_pipeline_submit(fdl_pipeline* p, const fdl_image* frames, int n, int* ticket) try {
  DeviceGuard _device_guard;
  if (!p || !frames || !ticket) return set_error(FDL_ERR_INVALID, "null argument");
  if (n <= 0 || n > p->cfg.max_batch) return set_error(FDL_ERR_INVALID, "batch size out of range (1..max_batch)");
  FDL_CUDA_TRY(cudaSetDevice(p->cfg.device));
  std::lock_guard<std::mutex> guard(p->mu);
  Lane* lane = nullptr;
  for (auto& l : p->lanes) if (!l.busy) { lane = &l; break; }
  if (!lane) return set_error(FDL_ERR_INVALID, "all pipeline lanes are in flight: collect a ticket first");
  for (int i = 0; i < n; ++i)
    if (frames[i].width != p->cfg.frame_width || frames[i].height != p->cfg.frame_height)
      return set_error(FDL_ERR_INVALID, "frame size differs from the pipeline configuration");
  // ---- copy-in (skipped for contiguous device frames and, in zero-copy mode, for pinned host frames)
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_h2d_start, lane->stream));
  int w, h;
  const uint8_t* fptr = nullptr;
  bool used_host = false;
  int rc = stage_frames(frames, n, &lane->frames, lane->stream, &w, &h, &fptr, p->cfg.zero_copy_host != 0, &used_host);
  if (rc) return rc;
  lane->jpeg_pending = false;
  return pipeline_enqueue(p, lane, n, fptr, used_host, frames[0].data, ticket);
} FDL_ABI_CATCH

int fdl_pipeline_submit_jpeg(fdl_pipeline* p, const uint8_t* const* data, const size_t* len, int n, int* ticket) try {
  DeviceGuard _device_guard;
  if (!p || !data || !len || !ticket) return set_error(FDL_ERR_INVALID, "null argument");
  if (n <= 0 || n > p->cfg.max_batch) return set_error(FDL_ERR_INVALID, "batch size out of range (1..max_batch)");
  FDL_CUDA_TRY(cudaSetDevice(p->cfg.device));
  std::lock_guard<std::mutex> guard(p->mu);
  Lane* lane = nullptr;
  for (auto& l : p->lanes) if (!l.busy) { lane = &l; break; }
  if (!lane) return set_error(FDL_ERR_INVALID, "all pipeline lanes are in flight: collect a ticket first");
  const int W = p->cfg.frame_width, H = p->cfg.frame_height;
  // convert_image_to_mat (utils.rs:8-21) for the whole batch, on the device: parse the headers here, copy the files compressed
  int rc = lane->jpeg.plan(data, len, n, W, H);
  if (rc) return rc;
  const size_t frame = (size_t)W * 3 * H;
  for (int i = 0; i < n; ++i) lane->jpeg.set_output(i, (long long)(frame * i), W * 3);
  FDL_CUDA_TRY(lane->frames.reserve(frame * (size_t)n));
  FDL_CUDA_TRY(cudaEventRecord(lane->ev_h2d_start, lane->stream));
  static const int poison_env = getenv("FDL_JPEG_POISON") ? atoi(getenv("FDL_JPEG_POISON")) : 0;     // tests: no stale pixel may be read
  if (poison_env) FDL_CUDA_TRY(cudaMemsetAsync(lane->frames.p, 0xA5, frame * (size_t)n, lane->stream));
  bool sparse = false;
  rc = lane->jpeg.enqueue(lane->frames.p, lane->stream, p->n_jpeg_rows > 0 ? p->jpeg_rows.p : nullptr, p->n_jpeg_rows, p->jpeg_rows_done.p, &sparse);
  if (rc) return rc;
  lane->jpeg_pending = true;
  return pipeline_enqueue(p, lane, n, lane->frames.p, false, nullptr, ticket, sparse ? &lane->jpeg : nullptr);
} FDL_ABI_CATCH

int fdl_pipeline_collect(fdl_pipeline* p, int ticket, fdl_frame_result* frame_results, fdl_face_result* face_results, int* n_out) try {
  DeviceGuard _device_guard;
  if (!p) return set_error(FDL_ERR_INVALID, "null argument");
  Lane* lane = nullptr;
  {
    std::lock_guard<std::mutex> guard(p->mu);
    for (auto& l : p->lanes) if (l.busy && l.ticket == ticket) { lane = &l; break; }
  }
  if (!lane) return set_error(FDL_ERR_INVALID, "unknown ticket");
  FDL_CUDA_TRY(cudaSetDevice(p->cfg.device));
  FDL_CUDA_TRY(cudaEventSynchronize(lane->ev_done));     // (a busy lane is not touched by submit)
  std::lock_guard<std::mutex> guard(p->mu);
  const int n = lane->n, F = n * p->cfg.max_faces;
  if (frame_results) std::memcpy(frame_results, lane->h_frames.p, (size_t)n * sizeof(fdl_frame_result));
  if (face_results && p->lmk) std::memcpy(face_results, lane->h_faces.p, (size_t)F * sizeof(fdl_face_result));
  if (n_out) *n_out = n;
  float ms = 0.f;
  cudaEventElapsedTime(&p->stage_ms[0], lane->ev_h2d_start, lane->ev_stage[0]);
  for (int i = 0; i < 8; ++i) {
    cudaEventElapsedTime(&ms, lane->ev_stage[i], lane->ev_stage[i + 1]);
    p->stage_ms[i + 1] = ms;
  }
  cudaEventElapsedTime(&p->stage_ms[9], lane->ev_stage[8], lane->ev_done);
  cudaEventElapsedTime(&p->last_device_ms, lane->ev_stage[0], lane->ev_stage[8]);
  lane->busy = false;
  if (lane->jpeg_pending) {
    lane->jpeg_pending = false;
    const int jrc = lane->jpeg.check_status();
    if (jrc) return jrc;
  }
  // the reference returns every detection (an unbounded Vec, face_detection.rs:267); a truncated record is reported, not hidden
  for (int i = 0; i < n; ++i)
    if (lane->h_frames.p[i].n_total_detections > FDL_MAX_DETECTIONS)
      return set_error(FDL_ERR_CAPACITY, "frame " + std::to_string(i) + " produced " + std::to_string(lane->h_frames.p[i].n_total_detections) +
                                             " detections: only the first FDL_MAX_DETECTIONS (32, in NMS order) are in its record");
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_pipeline_run(fdl_pipeline* p, const fdl_image* frames, int n, fdl_frame_result* frame_results, fdl_face_result* face_results) try {
  DeviceGuard _device_guard;
  int ticket = -1;
  int rc = fdl_pipeline_submit(p, frames, n, &ticket);
  if (rc) return rc;
  return fdl_pipeline_collect(p, ticket, frame_results, face_results, nullptr);
} FDL_ABI_CATCH

int fdl_letterbox_row_plan(int frame_width, int frame_height, int input_size, int32_t* row_pos, int32_t* info4) try {
  DeviceGuard _device_guard;
  if (frame_width <= 0 || frame_height <= 0 || input_size <= 0) return set_error(FDL_ERR_INVALID, "bad arguments");
  RowGather g;
  std::vector<int> rp;
  if (!plan_row_gather(frame_width, frame_height, input_size, &g, &rp)) {
    if (info4) info4[0] = info4[1] = info4[2] = info4[3] = 0;
    return 0;
  }
  if (row_pos) for (int i = 0; i < frame_height; ++i) row_pos[i] = rp[(size_t)i];
  if (info4) { info4[0] = g.rows_per_frame; info4[1] = g.period_src_rows; info4[2] = g.periods_per_frame; info4[3] = (int)g.fam.size(); }
  return 1;
} FDL_ABI_CATCH

float fdl_pipeline_last_device_ms(const fdl_pipeline* p) { return p ? p->last_device_ms : 0.f; }
int fdl_pipeline_stage_ms(const fdl_pipeline* p, float* out10) try {
  DeviceGuard _device_guard;
  if (!p || !out10) return set_error(FDL_ERR_INVALID, "null argument");
  // [0] H2D, [1] detector preprocess, [2] detector net, [3] SSD post, [4] face ROI + warp, [5] landmark net,
  // [6] landmark post + eye warp, [7] iris net, [8] iris post, [9] D2H
  for (int i = 0; i < kStages; ++i) out10[i] = p->stage_ms[i];
  return FDL_OK;
} FDL_ABI_CATCH

}  // extern "C"
