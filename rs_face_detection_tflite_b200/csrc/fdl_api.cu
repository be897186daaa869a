// fdl_api.cu -- the C ABI of include/fdl.h: handles, single-stage entry points, free functions.
// (The batched pipeline lives in pipeline.cu.)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "device_util.h"
#include "fdl_status.h"
#include "jpeg_decode.h"
#include "jpeg_parse.h"
#include "net.h"
#include "prepost_kernels.cuh"

namespace fdl {

static thread_local std::string g_last_error;
int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }
void clear_error() { g_last_error.clear(); }
unsigned long long launch_count_value();

// face_detection.rs:125-129
static const char* detector_file(int model) {
  switch (model) {
    case FDL_MODEL_FRONT_CAMERA: return "face_detection_front.tflite";
    case FDL_MODEL_BACK_CAMERA: return "face_detection_back.tflite";
    case FDL_MODEL_SHORT: return "face_detection_short_range.tflite";
    case FDL_MODEL_FULL: return "face_detection_full_range.tflite";
    case FDL_MODEL_FULL_SPARSE: return "face_detection_full_range_sparse.tflite";
    default: return nullptr;
  }
}

static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(FDL_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                       (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
  if (device < 0 || device >= n) return set_error(FDL_ERR_INVALID, "device index out of range");
  FDL_CUDA_TRY(cudaSetDevice(device));
  return FDL_OK;
}

}  // namespace fdl

using namespace fdl;

// ---------------------------------------------------------------------------------------------
struct fdl_net {
  Net* net = nullptr;
  bool owned = true;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
};

struct fdl_detector {
  fdl_net nh;
  int model = 0, device = 0;
  int S = 0, N = 0;
  SsdOptions opt{};
  cudaStream_t stream = nullptr;
  DevBuf<float> anchors;
  DevBuf<uint8_t> frames;
  DevBuf<I2TParams> params;
  DevBuf<fdl_rect> rois;
  PinBuf<fdl_rect> h_rois;
  DevBuf<fdl_detection> dets;
  DevBuf<int> counts;      // [2B]: clamped count, total count
  DevBuf<double> padding;
  DevBuf<float> raw_reg, raw_cls;
  DevBuf<int32_t> surv;    // [2*B*cap]
  DevBuf<int> nsurv;
};

struct fdl_landmark_model {
  fdl_net nh;
  int device = 0, S = 0;
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> frames;
  DevBuf<I2TParams> params;
  DevBuf<fdl_rect> rois;
  PinBuf<fdl_rect> h_rois;
  DevBuf<float> out;   // projected landmarks
  DevBuf<int> flags;
};

struct fdl_iris_model {
  fdl_net nh;
  int device = 0, S = 0;
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> frames;
  DevBuf<I2TParams> params;
  DevBuf<fdl_rect> rois;
  PinBuf<fdl_rect> h_rois;
  DevBuf<float> out;
};

namespace fdl { cudaError_t jpeg_debug_phases(long long out[8]); cudaError_t ws_trace_read(unsigned long long* out, int n);
cudaError_t chain_trace_read(unsigned long long* out, int n); }

struct fdl_frame {
  int device = 0;
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> buf;
  int w = 0, h = 0;
  JpegDecoder dec;
};

struct fdl_jpeg_decoder {
  int device = 0;
  cudaStream_t stream = nullptr;
  JpegDecoder dec;
  DevBuf<uint8_t> out;
};

extern "C" {

const char* fdl_last_error(void) { return g_last_error.c_str(); }
const char* fdl_version(void) { return "fdl-b200 0.1 (sm_100a)"; }
int fdl_device_count(void) try {
  DeviceGuard _device_guard;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
} FDL_ABI_CATCH
uint64_t fdl_launch_count(void) { return launch_count_value(); }

// ------------------------------------------------------------------------------------------ net
static int net_forward_host(Net* net, cudaStream_t stream, const float* in, int batch, float* const* outs, int n_outs) {
  if (!net || !in || batch <= 0 || !outs) return set_error(FDL_ERR_INVALID, "bad arguments");
  if (n_outs != net->num_outputs()) return set_error(FDL_ERR_INVALID, "wrong number of output buffers");
  std::string err;
  FDL_CUDA_TRY(cudaSetDevice(net->device() < 0 ? 0 : net->device()));
  if (!net->reserve(batch, &err)) return set_error(FDL_ERR_CUDA, err);
  TView iv = net->input_view(batch);
  FDL_CUDA_TRY(cudaMemcpyAsync(iv.p, in, (size_t)batch * net->in_elems() * sizeof(float), cudaMemcpyHostToDevice, stream));
  FDL_CUDA_TRY(net->forward(batch, stream));
  for (int i = 0; i < n_outs; ++i) {
    if (!outs[i]) continue;
    TView ov = net->output_view(i, batch);
    // outputs are root buffers: [batch, elems] contiguous
    FDL_CUDA_TRY(cudaMemcpyAsync(outs[i], ov.p, (size_t)batch * net->out_elems(i) * sizeof(float), cudaMemcpyDeviceToHost, stream));
  }
  FDL_CUDA_TRY(cudaStreamSynchronize(stream));
  return FDL_OK;
}

int fdl_net_create(const char* tflite_file, int device, fdl_net** out) try {
  DeviceGuard _device_guard;
  if (!tflite_file || !out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  if (device >= 0) { int rc = check_device(device); if (rc) return rc; }
  std::string err; int code = FDL_ERR_INTERNAL;
  Net* n = Net::create(tflite_file, device, &err, &code);
  if (!n) return set_error(code, err);
  fdl_net* h = new fdl_net();
  h->net = n; h->owned = true;
  if (device >= 0) {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete n; delete h; return set_error(FDL_ERR_CUDA, cudaGetErrorString(e)); }
    h->own_stream = true;
  }
  *out = h;
  return FDL_OK;
} FDL_ABI_CATCH
void fdl_net_destroy(fdl_net* h) {
  if (!h) return;
  DeviceGuard _device_guard;
  if (h->net && h->net->device() >= 0) cudaSetDevice(h->net->device());
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (h->owned) delete h->net;
  delete h;
}
fdl_net* fdl_detector_net(fdl_detector* d) { return d ? &d->nh : nullptr; }
fdl_net* fdl_landmark_net(fdl_landmark_model* m) { return m ? &m->nh : nullptr; }
fdl_net* fdl_iris_net(fdl_iris_model* m) { return m ? &m->nh : nullptr; }
int fdl_net_num_outputs(const fdl_net* h) { return h && h->net ? h->net->num_outputs() : 0; }
int64_t fdl_net_io_elems(const fdl_net* h, int i) try {
  DeviceGuard _device_guard;
  if (!h || !h->net) return 0;
  if (i < 0) return h->net->in_elems();
  return i < h->net->num_outputs() ? h->net->out_elems(i) : 0;
} FDL_ABI_CATCH
int fdl_net_forward(fdl_net* h, const float* in, int batch, float* const* outs, int n_outs) try {
  DeviceGuard _device_guard;
  if (!h || !h->net) return set_error(FDL_ERR_INVALID, "null handle");
  return net_forward_host(h->net, h->stream, in, batch, outs, n_outs);
} FDL_ABI_CATCH
int64_t fdl_net_describe(const fdl_net* h, char* buf, int64_t cap) try {
  DeviceGuard _device_guard;
  if (!h || !h->net) return 0;
  std::string s = h->net->plan().describe();
  if (buf && cap > 0) {
    size_t n = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int64_t)s.size() + 1;
} FDL_ABI_CATCH
int fdl_net_num_steps(const fdl_net* h) { return h && h->net ? (int)h->net->plan().steps.size() : 0; }
int fdl_net_set_mode(fdl_net* h, int mode) try {
  DeviceGuard _device_guard;
  if (!h || !h->net) return set_error(FDL_ERR_INVALID, "null handle");
  if (mode < 0 || mode > 2) return set_error(FDL_ERR_INVALID, "mode must be 0 (fp32), 1 (split-TF32 tensor cores) or 2 (same, serial BlazeBlock kernel only)");
  h->net->set_mode(mode);
  return FDL_OK;
} FDL_ABI_CATCH
int fdl_net_time_forward(fdl_net* h, const float* in, int batch, int iters, float* ms_per_pass) try {
  DeviceGuard _device_guard;
  if (!h || !h->net || batch <= 0 || iters <= 0 || !ms_per_pass) return set_error(FDL_ERR_INVALID, "bad arguments");
  Net* net = h->net;
  std::string err;
  if (net->device() < 0) return set_error(FDL_ERR_CUDA, "plan-only handle cannot run");
  FDL_CUDA_TRY(cudaSetDevice(net->device()));
  if (!net->reserve(batch, &err)) return set_error(FDL_ERR_CUDA, err);
  TView iv = net->input_view(batch);
  if (in) FDL_CUDA_TRY(cudaMemcpyAsync(iv.p, in, (size_t)batch * net->in_elems() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  FDL_CUDA_TRY(net->forward(batch, h->stream));  // warm-up
  EventSet evs;
  FDL_CUDA_TRY(evs.create(2));
  cudaEvent_t e0 = evs.ev[0], e1 = evs.ev[1];
  FDL_CUDA_TRY(cudaEventRecord(e0, h->stream));
  for (int i = 0; i < iters; ++i) FDL_CUDA_TRY(net->forward(batch, h->stream));
  FDL_CUDA_TRY(cudaEventRecord(e1, h->stream));
  FDL_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  FDL_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_pass = ms / (float)iters;
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_net_time_steps(fdl_net* h, const float* in, int batch, int iters, float* ms_per_step, int cap) try {
  DeviceGuard _device_guard;
  if (!h || !h->net || batch <= 0 || iters <= 0 || !ms_per_step) return set_error(FDL_ERR_INVALID, "bad arguments");
  Net* net = h->net;
  const int n = (int)net->plan().steps.size();
  if (cap < n) return set_error(FDL_ERR_CAPACITY, "ms_per_step too small");
  std::string err;
  if (net->device() < 0) return set_error(FDL_ERR_CUDA, "plan-only handle cannot run");
  FDL_CUDA_TRY(cudaSetDevice(net->device()));
  if (!net->reserve(batch, &err)) return set_error(FDL_ERR_CUDA, err);
  TView iv = net->input_view(batch);
  if (in) FDL_CUDA_TRY(cudaMemcpyAsync(iv.p, in, (size_t)batch * net->in_elems() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  FDL_CUDA_TRY(net->forward(batch, h->stream));  // warm-up
  EventSet evs;
  FDL_CUDA_TRY(evs.create(n + 1));
  cudaEvent_t* ev = evs.ev;
  for (int i = 0; i < n; ++i) ms_per_step[i] = 0.f;
  for (int it = 0; it < iters; ++it) {
    FDL_CUDA_TRY(net->forward(batch, h->stream, nullptr, nullptr, ev));
    FDL_CUDA_TRY(cudaEventSynchronize(ev[n]));
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      FDL_CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      ms_per_step[i] += ms / (float)iters;
    }
  }
  return FDL_OK;
} FDL_ABI_CATCH

// ------------------------------------------------------------------------------------- detector
int fdl_detector_create(int model, const char* model_dir, int device, fdl_detector** out) try {
  DeviceGuard _device_guard;
  if (!out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  const char* file = detector_file(model);
  SsdOptions opt;
  if (!file || !ssd_options_for(model, &opt)) return set_error(FDL_ERR_MODEL, "unsupported model type");   // face_detection.rs:184
  int rc = check_device(device);
  if (rc) return rc;
  std::string path = std::string(model_dir ? model_dir : "./models") + "/" + file;
  std::string err; int code = FDL_ERR_INTERNAL;
  Net* net = Net::create(path, device, &err, &code);
  if (!net) return set_error(code, err);
  fdl_detector* d = new fdl_detector();
  d->nh.net = net; d->nh.owned = true;
  d->model = model; d->device = device; d->opt = opt;
  d->S = net->plan().input.H;
  d->N = ssd_num_anchors(opt);
  auto bail = [&](int c, const std::string& m) { fdl_detector_destroy(d); return set_error(c, m); };
  if (net->num_outputs() != 2 || net->plan().input.W != d->S || d->S != opt.input_size || net->out_elems(0) != (int64_t)d->N * 16 ||
      net->out_elems(1) != d->N)
    return bail(FDL_ERR_MODEL, "incompatible model: expected [1,S,S,3] -> regressors [1,N,16], classificators [1,N,1]");
  cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = d->anchors.reserve((size_t)d->N * 2);
  if (e == cudaSuccess) e = launch_anchors(opt, d->anchors.p, d->N, d->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  if (e != cudaSuccess) return bail(FDL_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
  d->nh.stream = d->stream;
  *out = d;
  return FDL_OK;
} FDL_ABI_CATCH
void fdl_detector_destroy(fdl_detector* d) {
  if (!d) return;
  DeviceGuard _device_guard;
  cudaSetDevice(d->device);
  if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
  delete d->nh.net;
  delete d;
}
int fdl_detector_input_size(const fdl_detector* d) { return d ? d->S : 0; }
int fdl_detector_num_anchors(const fdl_detector* d) { return d ? d->N : 0; }
int fdl_detector_anchors(const fdl_detector* d, float* out_xy, int cap_anchors) try {
  DeviceGuard _device_guard;
  if (!d || !out_xy) return set_error(FDL_ERR_INVALID, "null argument");
  if (cap_anchors < d->N) return set_error(FDL_ERR_CAPACITY, "anchor buffer too small");
  FDL_CUDA_TRY(cudaSetDevice(d->device));
  FDL_CUDA_TRY(cudaMemcpy(out_xy, d->anchors.p, (size_t)d->N * 2 * sizeof(float), cudaMemcpyDeviceToHost));
  return FDL_OK;
} FDL_ABI_CATCH

// shared tail: SSD post-processing on device tensors + copy-out
static int detector_post(fdl_detector* d, const float* reg, long long reg_bs, const float* cls, long long cls_bs, int batch,
                         const I2TParams* params, const double* d_padding, fdl_detection* out, int cap, int* n_out,
                         int32_t* surv_anchor, int32_t* surv_cluster, int cap_surv, int* n_surv) {
  int cap_dev = cap < 1 ? 1 : (cap > d->N ? d->N : cap);
  FDL_CUDA_TRY(d->dets.reserve((size_t)batch * cap_dev));
  FDL_CUDA_TRY(d->counts.reserve((size_t)batch * 2));
  SsdPostArgs a;
  a.reg = reg; a.reg_bstride = reg_bs; a.cls = cls; a.cls_bstride = cls_bs;
  a.anchors = d->anchors.p; a.N = d->N; a.B = batch; a.scale = (float)d->S;
  a.params = params; a.padding4 = d_padding;
  a.det_base = reinterpret_cast<char*>(d->dets.p); a.det_stride = (long long)cap_dev * sizeof(fdl_detection);
  a.ndet_base = reinterpret_cast<char*>(d->counts.p); a.ndet_stride = sizeof(int);
  a.max_out = cap_dev; a.n_total = d->counts.p + batch;
  if (surv_anchor && surv_cluster && cap_surv > 0) {
    FDL_CUDA_TRY(d->surv.reserve((size_t)batch * cap_surv * 2));
    FDL_CUDA_TRY(d->nsurv.reserve((size_t)batch));
    a.surv_anchor = d->surv.p; a.surv_cluster = d->surv.p + (size_t)batch * cap_surv; a.cap_surv = cap_surv; a.n_surv = d->nsurv.p;
  }
  FDL_CUDA_TRY(launch_ssd_postprocess(a, d->stream));
  std::vector<int> counts((size_t)batch * 2);
  std::vector<fdl_detection> dets((size_t)batch * cap_dev);
  FDL_CUDA_TRY(cudaMemcpyAsync(counts.data(), d->counts.p, counts.size() * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(dets.data(), d->dets.p, dets.size() * sizeof(fdl_detection), cudaMemcpyDeviceToHost, d->stream));
  if (a.surv_anchor) {
    FDL_CUDA_TRY(cudaMemcpyAsync(surv_anchor, a.surv_anchor, (size_t)batch * cap_surv * sizeof(int32_t), cudaMemcpyDeviceToHost, d->stream));
    FDL_CUDA_TRY(cudaMemcpyAsync(surv_cluster, a.surv_cluster, (size_t)batch * cap_surv * sizeof(int32_t), cudaMemcpyDeviceToHost, d->stream));
    if (n_surv) FDL_CUDA_TRY(cudaMemcpyAsync(n_surv, d->nsurv.p, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  }
  FDL_CUDA_TRY(cudaStreamSynchronize(d->stream));
  bool overflow = false;
  for (int b = 0; b < batch; ++b) {
    int n = counts[b], total = counts[batch + b];
    if (total > cap) overflow = true;
    n_out[b] = total;
    if (out) for (int k = 0; k < n && k < cap; ++k) out[(size_t)b * cap + k] = dets[(size_t)b * cap_dev + k];
  }
  if (overflow) return set_error(FDL_ERR_CAPACITY, "more detections than the output capacity (n_out holds the required counts)");
  return FDL_OK;
}

static int detector_infer_impl(fdl_detector* d, const fdl_image* images, int batch, const fdl_rect* roi, fdl_detection* out, int cap,
                               int* n_out) {
  if (!d || !images || !n_out || batch <= 0 || cap < 0 || (cap > 0 && !out)) return set_error(FDL_ERR_INVALID, "bad arguments");
  FDL_CUDA_TRY(cudaSetDevice(d->device));
  int w, h;
  const uint8_t* fptr = nullptr;                  // device-resident, tightly packed images (an fdl_frame) are read in place
  int rc = stage_frames(images, batch, &d->frames, d->stream, &w, &h, &fptr);
  if (rc) return rc;
  Net* net = d->nh.net;
  std::string err;
  if (!net->reserve(batch, &err)) return set_error(FDL_ERR_CUDA, err);
  FDL_CUDA_TRY(d->params.reserve((size_t)batch));
  const fdl_rect* d_roi = nullptr;
  if (roi) {
    FDL_CUDA_TRY(d->rois.reserve((size_t)batch));
    // staged through a pinned slot: the copy is ordered on the stream, and the slot is free again when this call returns (every
    // infer ends with a stream synchronize) -- no extra synchronize per ROI
    FDL_CUDA_TRY(d->h_rois.reserve((size_t)batch));
    for (int b = 0; b < batch; ++b) d->h_rois.p[b] = *roi;
    FDL_CUDA_TRY(cudaMemcpyAsync(d->rois.p, d->h_rois.p, (size_t)batch * sizeof(fdl_rect), cudaMemcpyHostToDevice, d->stream));
    d_roi = d->rois.p;
  }
  // image_to_tensor(image, roi, (S,S), keep_aspect_ratio=true, (-1,1), flip=false)  face_detection.rs:219
  FDL_CUDA_TRY(launch_i2t_setup(d_roi, nullptr, nullptr, batch, w, h, d->S, d->S, 1, -1.0, 1.0, 0, d->params.p, nullptr, d->stream));
  TView iv = net->input_view(batch);
  FDL_CUDA_TRY(launch_i2t(fptr, (long long)w * 3 * h, (long long)w * 3, d->params.p, batch, d->S, d->S, iv.p, iv.bstride, nullptr,
                          nullptr, d->stream, 1, w));
  FDL_CUDA_TRY(net->forward(batch, d->stream));
  TView reg = net->output_view(0, batch), cls = net->output_view(1, batch);
  // a caller ROI of zero / negative size (or a singular transform) leaves an all-zero tensor behind: the reference errors there
  // (OpenCV throws; detection_letterbox_removal asserts its scales, transform.rs:121-122), so read the setup kernel's verdict back
  std::vector<I2TParams> hp((size_t)batch);
  FDL_CUDA_TRY(cudaMemcpyAsync(hp.data(), d->params.p, hp.size() * sizeof(I2TParams), cudaMemcpyDeviceToHost, d->stream));
  rc = detector_post(d, reg.p, reg.bstride, cls.p, cls.bstride, batch, d->params.p, nullptr, out, cap, n_out, nullptr, nullptr, 0, nullptr);
  for (int b = 0; b < batch; ++b) {
    if (!hp[(size_t)b].valid) return set_error(FDL_ERR_INVALID, "degenerate ROI: perspective transform is singular or empty");
    if (!(1.0 - (hp[(size_t)b].pad[0] + hp[(size_t)b].pad[2]) > 2.220446049250313e-16) || !(1.0 - (hp[(size_t)b].pad[1] + hp[(size_t)b].pad[3]) > 2.220446049250313e-16))
      return set_error(FDL_ERR_INVALID, "letterbox scale is too small");
  }
  return rc;
}

int fdl_detector_infer(fdl_detector* d, const fdl_image* image, const fdl_rect* roi, fdl_detection* out, int cap, int* n_out) try {
  DeviceGuard _device_guard;
  return detector_infer_impl(d, image, 1, roi, out, cap, n_out);
} FDL_ABI_CATCH
int fdl_detector_infer_batch(fdl_detector* d, const fdl_image* images, int batch, fdl_detection* out, int cap_per_image, int* n_out) try {
  DeviceGuard _device_guard;
  return detector_infer_impl(d, images, batch, nullptr, out, cap_per_image, n_out);
} FDL_ABI_CATCH
int fdl_detector_forward(fdl_detector* d, const float* in, int batch, float* regressors, float* classificators) try {
  DeviceGuard _device_guard;
  if (!d) return set_error(FDL_ERR_INVALID, "null handle");
  float* outs[2] = {regressors, classificators};
  return net_forward_host(d->nh.net, d->stream, in, batch, outs, 2);
} FDL_ABI_CATCH
int fdl_detector_postprocess(fdl_detector* d, const float* regressors, const float* classificators, int batch, const double* padding4,
                             fdl_detection* out, int cap_per_image, int* n_out, int32_t* survivor_anchor, int32_t* survivor_cluster,
                             int cap_surv, int* n_surv) try {
  DeviceGuard _device_guard;
  if (!d || !regressors || !classificators || batch <= 0 || !n_out) return set_error(FDL_ERR_INVALID, "bad arguments");
  FDL_CUDA_TRY(cudaSetDevice(d->device));
  FDL_CUDA_TRY(d->raw_reg.reserve((size_t)batch * d->N * 16));
  FDL_CUDA_TRY(d->raw_cls.reserve((size_t)batch * d->N));
  FDL_CUDA_TRY(cudaMemcpyAsync(d->raw_reg.p, regressors, (size_t)batch * d->N * 16 * sizeof(float), cudaMemcpyHostToDevice, d->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(d->raw_cls.p, classificators, (size_t)batch * d->N * sizeof(float), cudaMemcpyHostToDevice, d->stream));
  const double* d_pad = nullptr;
  if (padding4) {
    // detection_letterbox_removal asserts both scales > f64::EPSILON (transform.rs:121-122)
    for (int b = 0; b < batch; ++b) {
      const double* q = padding4 + 4 * (size_t)b;
      if (!(1.0 - (q[0] + q[2]) > 2.220446049250313e-16)) return set_error(FDL_ERR_INVALID, "Horizontal scale is too small");
      if (!(1.0 - (q[1] + q[3]) > 2.220446049250313e-16)) return set_error(FDL_ERR_INVALID, "Vertical scale is too small");
    }
    FDL_CUDA_TRY(d->padding.reserve((size_t)batch * 4));
    FDL_CUDA_TRY(cudaMemcpyAsync(d->padding.p, padding4, (size_t)batch * 4 * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    d_pad = d->padding.p;
  }
  return detector_post(d, d->raw_reg.p, (long long)d->N * 16, d->raw_cls.p, d->N, batch, nullptr, d_pad, out, cap_per_image, n_out,
                       survivor_anchor, survivor_cluster, cap_surv, n_surv);
} FDL_ABI_CATCH

// ------------------------------------------------------------------------------------- landmark
int fdl_landmark_create(const char* model_file, int device, fdl_landmark_model** out) try {
  DeviceGuard _device_guard;
  if (!out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  std::string err; int code = FDL_ERR_INTERNAL;
  Net* net = Net::create(model_file ? model_file : "./models/face_landmark.tflite", device, &err, &code);   // face_landmark.rs:211-215
  if (!net) return set_error(code, err);
  fdl_landmark_model* m = new fdl_landmark_model();
  m->nh.net = net; m->device = device; m->S = net->plan().input.H;
  // face_landmark.rs:244-247: last output dim must hold NUM_DIMS * NUM_LANDMARKS values
  if (net->num_outputs() != 2 || net->out_elems(0) < 3 * FDL_NUM_FACE_LANDMARKS || net->plan().input.W != m->S) {
    fdl_landmark_destroy(m);
    return set_error(FDL_ERR_MODEL, "incompatible model: landmark output smaller than 1404 values");
  }
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { fdl_landmark_destroy(m); return set_error(FDL_ERR_CUDA, cudaGetErrorString(e)); }
  m->nh.stream = m->stream;
  *out = m;
  return FDL_OK;
} FDL_ABI_CATCH
void fdl_landmark_destroy(fdl_landmark_model* m) {
  if (!m) return;
  DeviceGuard _device_guard;
  cudaSetDevice(m->device);
  if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
  delete m->nh.net;
  delete m;
}

__global__ void flag_gate_kernel(const float* flag, int* has) {
  // face_landmark.rs:292-296
  *has = !(sigmoid_f32(*flag) <= 0.5f) ? 1 : 0;
}

int fdl_landmark_infer(fdl_landmark_model* m, const fdl_image* image, const fdl_rect* roi, fdl_landmark* out, int* n_out,
                       float* face_flag_logit) try {
  DeviceGuard _device_guard;
  if (!m || !image || !out || !n_out) return set_error(FDL_ERR_INVALID, "bad arguments");
  FDL_CUDA_TRY(cudaSetDevice(m->device));
  int w, h;
  const uint8_t* fptr = nullptr;                  // a device-resident image (an fdl_frame) is read in place
  int rc = stage_frames(image, 1, &m->frames, m->stream, &w, &h, &fptr);
  if (rc) return rc;
  Net* net = m->nh.net;
  std::string err;
  if (!net->reserve(1, &err)) return set_error(FDL_ERR_CUDA, err);
  FDL_CUDA_TRY(m->params.reserve(1));
  FDL_CUDA_TRY(m->rois.reserve(1));
  FDL_CUDA_TRY(m->out.reserve(3 * FDL_NUM_FACE_LANDMARKS));
  FDL_CUDA_TRY(m->flags.reserve(2));
  const fdl_rect* d_roi = nullptr;
  if (roi) {
    FDL_CUDA_TRY(m->h_rois.reserve(1));
    m->h_rois.p[0] = *roi;                       // pinned slot, free again when this call returns
    FDL_CUDA_TRY(cudaMemcpyAsync(m->rois.p, m->h_rois.p, sizeof(fdl_rect), cudaMemcpyHostToDevice, m->stream));
    d_roi = m->rois.p;
  }
  // image_to_tensor(image, roi, (S,S), keep_aspect_ratio=false, (0,1), flip=false)  face_landmark.rs:250
  FDL_CUDA_TRY(launch_i2t_setup(d_roi, nullptr, nullptr, 1, w, h, m->S, m->S, 0, 0.0, 1.0, 0, m->params.p, nullptr, m->stream));
  TView iv = net->input_view(1);
  FDL_CUDA_TRY(launch_i2t(fptr, (long long)w * 3 * h, (long long)w * 3, m->params.p, 1, m->S, m->S, iv.p, iv.bstride, nullptr, nullptr,
                          m->stream));
  FDL_CUDA_TRY(net->forward(1, m->stream));
  TView raw = net->output_view(0, 1), flag = net->output_view(1, 1);
  // the reference takes the LAST element of the flag tensor (face_landmark.rs:292-293)
  const float* flag_last = flag.p + net->out_elems(1) - 1;
  flag_gate_kernel<<<1, 1, 0, m->stream>>>(flag_last, m->flags.p);
  FDL_CUDA_TRY(cudaGetLastError());
  const double* d_pad = reinterpret_cast<const double*>(reinterpret_cast<const char*>(m->params.p) + offsetof(I2TParams, pad));
  FDL_CUDA_TRY(launch_project(raw.p, FDL_NUM_FACE_LANDMARKS, m->S, m->S, w, h, d_pad, d_roi, 0, m->out.p, m->stream));
  std::vector<float> pts(3 * FDL_NUM_FACE_LANDMARKS);
  int has = 0; float logit = 0.f; I2TParams P;
  FDL_CUDA_TRY(cudaMemcpyAsync(pts.data(), m->out.p, pts.size() * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(&has, m->flags.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(&logit, flag_last, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(&P, m->params.p, sizeof(I2TParams), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (!P.valid) return set_error(FDL_ERR_INVALID, "degenerate ROI: perspective transform is singular or empty");
  if (face_flag_logit) *face_flag_logit = logit;
  if (!has) { *n_out = 0; return FDL_OK; }
  for (int k = 0; k < FDL_NUM_FACE_LANDMARKS; ++k) { out[k].x = pts[3 * k]; out[k].y = pts[3 * k + 1]; out[k].z = pts[3 * k + 2]; }
  *n_out = FDL_NUM_FACE_LANDMARKS;
  return FDL_OK;
} FDL_ABI_CATCH
int fdl_landmark_forward(fdl_landmark_model* m, const float* in, int batch, float* landmarks, float* flag) try {
  DeviceGuard _device_guard;
  if (!m) return set_error(FDL_ERR_INVALID, "null handle");
  float* outs[2] = {landmarks, flag};
  return net_forward_host(m->nh.net, m->stream, in, batch, outs, 2);
} FDL_ABI_CATCH

// ----------------------------------------------------------------------------------------- iris
int fdl_iris_create(const char* model_file, int device, fdl_iris_model** out) try {
  DeviceGuard _device_guard;
  if (!out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  std::string err; int code = FDL_ERR_INTERNAL;
  Net* net = Net::create(model_file ? model_file : "./models/iris_landmark.tflite", device, &err, &code);   // iris_landmark.rs:145-149
  if (!net) return set_error(code, err);
  fdl_iris_model* m = new fdl_iris_model();
  m->nh.net = net; m->device = device; m->S = net->plan().input.H;
  // iris_landmark.rs:172-184
  if (net->num_outputs() != 2 || net->out_elems(0) != 3 * FDL_NUM_EYE_CONTOUR || net->out_elems(1) != 3 * FDL_NUM_IRIS ||
      net->plan().input.W != m->S) {
    fdl_iris_destroy(m);
    return set_error(FDL_ERR_MODEL, "incompatible model: expected outputs [1,213] and [1,15]");
  }
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { fdl_iris_destroy(m); return set_error(FDL_ERR_CUDA, cudaGetErrorString(e)); }
  m->nh.stream = m->stream;
  *out = m;
  return FDL_OK;
} FDL_ABI_CATCH
void fdl_iris_destroy(fdl_iris_model* m) {
  if (!m) return;
  DeviceGuard _device_guard;
  cudaSetDevice(m->device);
  if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
  delete m->nh.net;
  delete m;
}
int fdl_iris_infer(fdl_iris_model* m, const fdl_image* image, const fdl_rect* roi, int is_right_eye, fdl_landmark* contour,
                   fdl_landmark* iris) try {
  DeviceGuard _device_guard;
  if (!m || !image || !contour || !iris) return set_error(FDL_ERR_INVALID, "bad arguments");
  FDL_CUDA_TRY(cudaSetDevice(m->device));
  int w, h;
  const uint8_t* fptr = nullptr;                  // a device-resident image (an fdl_frame) is read in place
  int rc = stage_frames(image, 1, &m->frames, m->stream, &w, &h, &fptr);
  if (rc) return rc;
  Net* net = m->nh.net;
  std::string err;
  if (!net->reserve(1, &err)) return set_error(FDL_ERR_CUDA, err);
  const int K = FDL_NUM_EYE_CONTOUR + FDL_NUM_IRIS;
  FDL_CUDA_TRY(m->params.reserve(1));
  FDL_CUDA_TRY(m->rois.reserve(1));
  FDL_CUDA_TRY(m->out.reserve(3 * K));
  const fdl_rect* d_roi = nullptr;
  if (roi) {
    FDL_CUDA_TRY(m->h_rois.reserve(1));
    m->h_rois.p[0] = *roi;                       // pinned slot, free again when this call returns
    FDL_CUDA_TRY(cudaMemcpyAsync(m->rois.p, m->h_rois.p, sizeof(fdl_rect), cudaMemcpyHostToDevice, m->stream));
    d_roi = m->rois.p;
  }
  // image_to_tensor(image, roi, (S,S), keep_aspect_ratio=true, (0,1), flip=is_right_eye)  iris_landmark.rs:188-189
  FDL_CUDA_TRY(launch_i2t_setup(d_roi, nullptr, nullptr, 1, w, h, m->S, m->S, 1, 0.0, 1.0, is_right_eye ? 1 : 0, m->params.p, nullptr,
                                m->stream));
  TView iv = net->input_view(1);
  FDL_CUDA_TRY(launch_i2t(fptr, (long long)w * 3 * h, (long long)w * 3, m->params.p, 1, m->S, m->S, iv.p, iv.bstride, nullptr, nullptr,
                          m->stream));
  FDL_CUDA_TRY(net->forward(1, m->stream));
  TView eye = net->output_view(0, 1), ir = net->output_view(1, 1);
  const double* d_pad = reinterpret_cast<const double*>(reinterpret_cast<const char*>(m->params.p) + offsetof(I2TParams, pad));
  FDL_CUDA_TRY(launch_project(eye.p, FDL_NUM_EYE_CONTOUR, m->S, m->S, w, h, d_pad, d_roi, is_right_eye ? 1 : 0, m->out.p, m->stream));
  FDL_CUDA_TRY(launch_project(ir.p, FDL_NUM_IRIS, m->S, m->S, w, h, d_pad, d_roi, is_right_eye ? 1 : 0, m->out.p + 3 * FDL_NUM_EYE_CONTOUR,
                              m->stream));
  std::vector<float> pts(3 * K);
  I2TParams P;
  FDL_CUDA_TRY(cudaMemcpyAsync(pts.data(), m->out.p, pts.size() * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaMemcpyAsync(&P, m->params.p, sizeof(I2TParams), cudaMemcpyDeviceToHost, m->stream));
  FDL_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (!P.valid) return set_error(FDL_ERR_INVALID, "degenerate ROI: perspective transform is singular or empty");
  for (int k = 0; k < FDL_NUM_EYE_CONTOUR; ++k) { contour[k].x = pts[3 * k]; contour[k].y = pts[3 * k + 1]; contour[k].z = pts[3 * k + 2]; }
  for (int k = 0; k < FDL_NUM_IRIS; ++k) {
    const float* q = &pts[3 * (FDL_NUM_EYE_CONTOUR + k)];
    iris[k].x = q[0]; iris[k].y = q[1]; iris[k].z = q[2];
  }
  return FDL_OK;
} FDL_ABI_CATCH
int fdl_iris_forward(fdl_iris_model* m, const float* in, int batch, float* contours, float* iris) try {
  DeviceGuard _device_guard;
  if (!m) return set_error(FDL_ERR_INVALID, "null handle");
  float* outs[2] = {contours, iris};
  return net_forward_host(m->nh.net, m->stream, in, batch, outs, 2);
} FDL_ABI_CATCH

// ------------------------------------------------------------------------------- free functions
// Small per-call scratch; the work itself is one thread on the device (same code as the pipeline).
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int rc(size_t n) {
    if (n <= cap) return FDL_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    FDL_CUDA_TRY(cudaMalloc(&p, n));
    cap = n;
    return FDL_OK;
  }
};

int fdl_face_detection_to_roi(int device, const fdl_detection* det, int image_width, int image_height, int size_mode, fdl_rect* out) try {
  DeviceGuard _device_guard;
  if (!det || !out) return set_error(FDL_ERR_INVALID, "null argument");
  if (size_mode < FDL_SIZE_MODE_NONE || size_mode > FDL_SIZE_MODE_SQUARE_SHORT) return set_error(FDL_ERR_INVALID, "bad size mode");
  int rc = check_device(device);
  if (rc) return rc;
  char* buf = nullptr;
  FDL_CUDA_TRY(cudaMalloc(&buf, 256));
  fdl_detection* d_det = reinterpret_cast<fdl_detection*>(buf);
  fdl_rect* d_rect = reinterpret_cast<fdl_rect*>(buf + 128);
  int* d_ok = reinterpret_cast<int*>(buf + 192);
  cudaError_t e = cudaMemcpy(d_det, det, sizeof(fdl_detection), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_face_detection_to_roi(d_det, image_width, image_height, size_mode, d_rect, d_ok, 0);
  int ok = 0;
  if (e == cudaSuccess) e = cudaMemcpy(out, d_rect, sizeof(fdl_rect), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (e != cudaSuccess) return set_error(FDL_ERR_CUDA, cudaGetErrorString(e));
  if (!ok) return set_error(FDL_ERR_INVALID, "bbox must be normalized");   // transform.rs:52
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_iris_roi_from_face_landmarks(int device, const fdl_landmark* landmarks, int n, int image_width, int image_height, fdl_rect* left,
                                     fdl_rect* right) try {
  DeviceGuard _device_guard;
  if (!landmarks || !left || !right) return set_error(FDL_ERR_INVALID, "null argument");
  if (n <= 362) return set_error(FDL_ERR_INVALID, "landmarks must contain the 468 face landmarks (indices 33,133,362,263 are read)");
  int rc = check_device(device);
  if (rc) return rc;
  const int idx[4] = {33, 133, 362, 263};   // iris_landmark.rs:29-35
  double xy[8];
  for (int k = 0; k < 4; ++k) { xy[2 * k] = landmarks[idx[k]].x; xy[2 * k + 1] = landmarks[idx[k]].y; }
  char* buf = nullptr;
  FDL_CUDA_TRY(cudaMalloc(&buf, 256));
  double* d_xy = reinterpret_cast<double*>(buf);
  fdl_rect* d_rect = reinterpret_cast<fdl_rect*>(buf + 64);
  int* d_ok = reinterpret_cast<int*>(buf + 192);
  cudaError_t e = cudaMemcpy(d_xy, xy, sizeof(xy), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_eye_rois(d_xy, image_width, image_height, d_rect, d_ok, 0);
  fdl_rect r[2]; int ok = 0;
  if (e == cudaSuccess) e = cudaMemcpy(r, d_rect, sizeof(r), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (e != cudaSuccess) return set_error(FDL_ERR_CUDA, cudaGetErrorString(e));
  if (!ok) return set_error(FDL_ERR_INVALID, "bbox must be normalized");
  *left = r[0]; *right = r[1];
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_update_face_landmarks_with_iris_results(int device, const fdl_landmark* face_landmarks, int n, const fdl_landmark* left_contour, int n_left,
                                                const fdl_landmark* right_contour, int n_right, fdl_landmark* refined) try {
  DeviceGuard _device_guard;
  if (!face_landmarks || !refined || (n_left > 0 && !left_contour) || (n_right > 0 && !right_contour)) return set_error(FDL_ERR_INVALID, "null argument");
  if (n != FDL_NUM_FACE_LANDMARKS) return set_error(FDL_ERR_INVALID, "unexpected number of items in face_landmarks");   // iris_landmark.rs:383-385
  if (n_left < 0 || n_left > FDL_NUM_EYE_CONTOUR || n_right < 0 || n_right > FDL_NUM_EYE_CONTOUR)
    return set_error(FDL_ERR_INVALID, "an eye contour has at most 71 points");
  int rc = check_device(device);
  if (rc) return rc;
  static_assert(sizeof(fdl_landmark) == 3 * sizeof(double), "fdl_landmark is three packed doubles");
  DevBuf<double> d;
  const size_t nf = 3 * FDL_NUM_FACE_LANDMARKS, ne = 3 * FDL_NUM_EYE_CONTOUR;
  FDL_CUDA_TRY(d.reserve(2 * nf + 2 * ne));
  double *d_face = d.p, *d_left = d.p + nf, *d_right = d_left + ne, *d_out = d_right + ne;
  FDL_CUDA_TRY(cudaMemcpy(d_face, face_landmarks, nf * sizeof(double), cudaMemcpyHostToDevice));
  if (n_left) FDL_CUDA_TRY(cudaMemcpy(d_left, left_contour, (size_t)3 * n_left * sizeof(double), cudaMemcpyHostToDevice));
  if (n_right) FDL_CUDA_TRY(cudaMemcpy(d_right, right_contour, (size_t)3 * n_right * sizeof(double), cudaMemcpyHostToDevice));
  FDL_CUDA_TRY(launch_refine_landmarks(d_face, d_left, n_left, d_right, n_right, d_out, 0));
  FDL_CUDA_TRY(cudaMemcpy(refined, d_out, nf * sizeof(double), cudaMemcpyDeviceToHost));
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_eye_to_face_landmark_index(int is_right_eye, int32_t* out71) try {
  DeviceGuard _device_guard;
  if (!out71) return set_error(FDL_ERR_INVALID, "null argument");
  for (int k = 0; k < FDL_NUM_EYE_CONTOUR; ++k) out71[k] = eye_to_face_landmark_index(is_right_eye ? 1 : 0, k);
  return FDL_OK;
} FDL_ABI_CATCH

static int iris_metrics(int device, const fdl_landmark* iris, int n, double focal_length_mm, double iris_size_px, int w, int h, double* out2) {
  if (!iris) return set_error(FDL_ERR_INVALID, "null argument");
  if (n < FDL_NUM_IRIS) return set_error(FDL_ERR_INVALID, "iris landmarks must hold the 5 IrisIndex points");   // the reference indexes [0..4] and panics
  if (w <= 0 || h <= 0) return set_error(FDL_ERR_INVALID, "image size must be positive");
  int rc = check_device(device);
  if (rc) return rc;
  DevBuf<double> d;
  FDL_CUDA_TRY(d.reserve(3 * FDL_NUM_IRIS + 2));
  FDL_CUDA_TRY(cudaMemcpy(d.p, iris, 3 * FDL_NUM_IRIS * sizeof(double), cudaMemcpyHostToDevice));
  FDL_CUDA_TRY(launch_iris_metrics(d.p, w, h, focal_length_mm, iris_size_px, d.p + 3 * FDL_NUM_IRIS, 0));
  FDL_CUDA_TRY(cudaMemcpy(out2, d.p + 3 * FDL_NUM_IRIS, 2 * sizeof(double), cudaMemcpyDeviceToHost));
  return FDL_OK;
}

int fdl_iris_diameter(int device, const fdl_landmark* iris, int n, int image_width, int image_height, double* diameter_px) try {
  DeviceGuard _device_guard;
  if (!diameter_px) return set_error(FDL_ERR_INVALID, "null argument");
  double o[2];
  int rc = iris_metrics(device, iris, n, 0.0, 0.0, image_width, image_height, o);
  if (rc) return rc;
  *diameter_px = o[0];
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_iris_depth(int device, const fdl_landmark* iris, int n, double focal_length_mm, double iris_size_px, int image_width, int image_height,
                   double* depth_mm) try {
  DeviceGuard _device_guard;
  if (!depth_mm) return set_error(FDL_ERR_INVALID, "null argument");
  if (!(iris_size_px > 0.0)) return set_error(FDL_ERR_INVALID, "iris_size_px must be positive");
  double o[2];
  int rc = iris_metrics(device, iris, n, focal_length_mm, iris_size_px, image_width, image_height, o);
  if (rc) return rc;
  *depth_mm = o[1];
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_image_to_tensor(int device, const fdl_image* image, const fdl_rect* roi, int out_w, int out_h, int keep_aspect_ratio,
                        double range_min, double range_max, int flip_horizontal, float* out_tensor, uint8_t* out_u8, double* padding4) try {
  DeviceGuard _device_guard;
  if (!image || !out_tensor || out_w <= 0 || out_h <= 0) return set_error(FDL_ERR_INVALID, "bad arguments");
  if (keep_aspect_ratio && out_w != out_h)
    return set_error(FDL_ERR_INVALID, "keep_aspect_ratio requires a square output (the reference divides the sizes as integers, transform.rs:240)");
  int rc = check_device(device);
  if (rc) return rc;
  DevBuf<uint8_t> frames; DevBuf<I2TParams> params; DevBuf<fdl_rect> rois; DevBuf<float> out; DevBuf<uint8_t> u8;
  int w, h;
  rc = stage_frames(image, 1, &frames, 0, &w, &h);
  if (rc) return rc;
  FDL_CUDA_TRY(params.reserve(1));
  FDL_CUDA_TRY(out.reserve((size_t)out_w * out_h * 3));
  FDL_CUDA_TRY(u8.reserve((size_t)out_w * out_h * 3));
  const fdl_rect* d_roi = nullptr;
  if (roi) {
    FDL_CUDA_TRY(rois.reserve(1));
    FDL_CUDA_TRY(cudaMemcpy(rois.p, roi, sizeof(fdl_rect), cudaMemcpyHostToDevice));
    d_roi = rois.p;
  }
  FDL_CUDA_TRY(launch_i2t_setup(d_roi, nullptr, nullptr, 1, w, h, out_w, out_h, keep_aspect_ratio, range_min, range_max, flip_horizontal ? 1 : 0,
                                params.p, nullptr, 0));
  FDL_CUDA_TRY(launch_i2t(frames.p, (long long)w * 3 * h, (long long)w * 3, params.p, 1, out_w, out_h, out.p, (long long)out_w * out_h * 3, u8.p,
                          nullptr, 0));
  I2TParams P;
  FDL_CUDA_TRY(cudaMemcpy(out_tensor, out.p, (size_t)out_w * out_h * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  if (out_u8) FDL_CUDA_TRY(cudaMemcpy(out_u8, u8.p, (size_t)out_w * out_h * 3, cudaMemcpyDeviceToHost));
  FDL_CUDA_TRY(cudaMemcpy(&P, params.p, sizeof(I2TParams), cudaMemcpyDeviceToHost));
  if (!P.valid) return set_error(FDL_ERR_INVALID, "degenerate ROI: perspective transform is singular or empty");
  if (padding4) for (int i = 0; i < 4; ++i) padding4[i] = P.pad[i];
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_project_landmarks(int device, const float* raw, int n, int tensor_w, int tensor_h, int image_w, int image_h, const double* padding4,
                          const fdl_rect* roi, int flip_horizontal, fdl_landmark* out) try {
  DeviceGuard _device_guard;
  if (!raw || !out || n <= 0 || !padding4) return set_error(FDL_ERR_INVALID, "bad arguments");
  int rc = check_device(device);
  if (rc) return rc;
  DevBuf<float> d_raw, d_out; DevBuf<double> d_pad; DevBuf<fdl_rect> d_roi;
  FDL_CUDA_TRY(d_raw.reserve((size_t)3 * n));
  FDL_CUDA_TRY(d_out.reserve((size_t)3 * n));
  FDL_CUDA_TRY(d_pad.reserve(4));
  FDL_CUDA_TRY(cudaMemcpy(d_raw.p, raw, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice));
  FDL_CUDA_TRY(cudaMemcpy(d_pad.p, padding4, 4 * sizeof(double), cudaMemcpyHostToDevice));
  if (roi) {
    FDL_CUDA_TRY(d_roi.reserve(1));
    FDL_CUDA_TRY(cudaMemcpy(d_roi.p, roi, sizeof(fdl_rect), cudaMemcpyHostToDevice));
  }
  FDL_CUDA_TRY(launch_project(d_raw.p, n, tensor_w, tensor_h, image_w, image_h, d_pad.p, roi ? d_roi.p : nullptr, flip_horizontal, d_out.p, 0));
  std::vector<float> pts((size_t)3 * n);
  FDL_CUDA_TRY(cudaMemcpy(pts.data(), d_out.p, pts.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int k = 0; k < n; ++k) { out[k].x = pts[3 * k]; out[k].y = pts[3 * k + 1]; out[k].z = pts[3 * k + 2]; }
  return FDL_OK;
} FDL_ABI_CATCH

// ---------------------------------------------------------------------------------- frame ingest
int fdl_jpeg_info(const uint8_t* data, size_t len, int* width, int* height, int* components) try {
  if (!data) return set_error(FDL_ERR_INVALID, "null argument");
  JpegHeader hd;
  std::string msg;
  if (!jpeg_parse_header(data, len, &hd, &msg)) return set_error(FDL_ERR_INVALID, msg);
  if (width) *width = hd.width;
  if (height) *height = hd.height;
  if (components) *components = hd.ncomp;
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_jpeg_decoder_create(int device, fdl_jpeg_decoder** out) try {
  DeviceGuard _device_guard;
  if (!out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  fdl_jpeg_decoder* d = new fdl_jpeg_decoder();
  d->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete d; return set_error(FDL_ERR_CUDA, cudaGetErrorString(e)); }
  *out = d;
  return FDL_OK;
} FDL_ABI_CATCH

void fdl_jpeg_decoder_destroy(fdl_jpeg_decoder* d) {
  if (!d) return;
  DeviceGuard _device_guard;
  cudaSetDevice(d->device);
  if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
  delete d;
}

int fdl_jpeg_decode(fdl_jpeg_decoder* d, const uint8_t* const* data, const size_t* len, int n, uint8_t* out, size_t cap, int out_mem,
                    int64_t* offsets, int32_t* widths, int32_t* heights) try {
  DeviceGuard _device_guard;
  if (!d) return set_error(FDL_ERR_INVALID, "null handle");
  FDL_CUDA_TRY(cudaSetDevice(d->device));
  int rc = d->dec.plan(data, len, n, 0, 0);
  if (rc) return rc;
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    const int w = d->dec.width(i), h = d->dec.height(i);
    if (offsets) offsets[i] = (int64_t)total;
    if (widths) widths[i] = w;
    if (heights) heights[i] = h;
    d->dec.set_output(i, (long long)total, w * 3);
    total += (size_t)w * 3 * (size_t)h;
    total = (total + 3) & ~size_t(3);            // every image starts 4-byte aligned (word stores of the colour kernel)
  }
  if (!out || cap < total) return set_error(FDL_ERR_CAPACITY, "output buffer too small: " + std::to_string(total) + " bytes needed");
  uint8_t* dst = out;
  if (out_mem != FDL_MEM_DEVICE) { FDL_CUDA_TRY(d->out.reserve(total)); dst = d->out.p; }
  rc = d->dec.enqueue(dst, d->stream);
  if (rc) return rc;
  if (out_mem != FDL_MEM_DEVICE) FDL_CUDA_TRY(cudaMemcpyAsync(out, d->out.p, total, cudaMemcpyDeviceToHost, d->stream));
  FDL_CUDA_TRY(cudaStreamSynchronize(d->stream));
  return d->dec.check_status();
} FDL_ABI_CATCH

// not in fdl.h: debugging aid (phase timestamps of CTA 0 of the last entropy-stage launch on the current device)
FDL_API int fdl_debug_jpeg_phases(long long* out8) {
  if (!out8) return FDL_ERR_INVALID;
  return fdl::jpeg_debug_phases(out8) == cudaSuccess ? FDL_OK : FDL_ERR_CUDA;
}

// not in fdl.h: timeline of block_ws_kernel's first CTAs (variant build with -DFDL_WS_TRACE only; FDL_ERR_INVALID otherwise)
FDL_API int fdl_debug_ws_trace(unsigned long long* out, int n) {
  if (!out) return FDL_ERR_INVALID;
  return fdl::ws_trace_read(out, n) == cudaSuccess ? FDL_OK : FDL_ERR_INVALID;
}

// not in fdl.h: per-op timeline of the tail-chain kernel's CTA 0 (variant build with -DFDL_WS_TRACE only)
FDL_API int fdl_debug_chain_trace(unsigned long long* out, int n) {
  if (!out) return FDL_ERR_INVALID;
  return fdl::chain_trace_read(out, n) == cudaSuccess ? FDL_OK : FDL_ERR_INVALID;
}

int fdl_decode_jpeg(int device, const uint8_t* data, size_t len, uint8_t* out_rgb, size_t cap, int* width, int* height) try {
  DeviceGuard _device_guard;
  int w = 0, h = 0;
  int rc = fdl_jpeg_info(data, len, &w, &h, nullptr);
  if (rc) return rc;
  if (width) *width = w;
  if (height) *height = h;
  if (!out_rgb || cap < (size_t)w * 3 * (size_t)h) return set_error(FDL_ERR_CAPACITY, "output buffer too small for " + std::to_string(w) + "x" + std::to_string(h) + " RGB");
  // convert_image_to_mat is a free function in the reference (utils.rs:8): one cached decoder per device serves it
  static std::mutex mu;
  static fdl_jpeg_decoder* cache[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (device < 0 || device >= 64) return set_error(FDL_ERR_INVALID, "device index out of range");
  if (!cache[device]) { rc = fdl_jpeg_decoder_create(device, &cache[device]); if (rc) return rc; }
  return fdl_jpeg_decode(cache[device], &data, &len, 1, out_rgb, cap, FDL_MEM_HOST, nullptr, nullptr, nullptr);
} FDL_ABI_CATCH

// ---------------------------------------------------------------------------------- fdl_frame
int fdl_frame_create(int device, fdl_frame** out) try {
  DeviceGuard _device_guard;
  if (!out) return set_error(FDL_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = check_device(device);
  if (rc) return rc;
  fdl_frame* f = new fdl_frame();
  f->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete f; return set_error(FDL_ERR_CUDA, cudaGetErrorString(e)); }
  *out = f;
  return FDL_OK;
} FDL_ABI_CATCH

void fdl_frame_destroy(fdl_frame* f) {
  if (!f) return;
  DeviceGuard _device_guard;
  cudaSetDevice(f->device);
  if (f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
  delete f;
}

int fdl_frame_upload(fdl_frame* f, const fdl_image* image) try {
  DeviceGuard _device_guard;
  if (!f || !image) return set_error(FDL_ERR_INVALID, "null argument");
  FDL_CUDA_TRY(cudaSetDevice(f->device));
  f->w = f->h = 0;
  int w, h;
  int rc = stage_frames(image, 1, &f->buf, f->stream, &w, &h);
  if (rc) return rc;
  FDL_CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->w = w; f->h = h;
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_frame_upload_jpeg(fdl_frame* f, const uint8_t* data, size_t len) try {
  DeviceGuard _device_guard;
  if (!f || !data) return set_error(FDL_ERR_INVALID, "null argument");
  FDL_CUDA_TRY(cudaSetDevice(f->device));
  f->w = f->h = 0;
  int rc = f->dec.plan(&data, &len, 1, 0, 0);
  if (rc) return rc;
  const int w = f->dec.width(0), h = f->dec.height(0);
  f->dec.set_output(0, 0, w * 3);
  FDL_CUDA_TRY(f->buf.reserve((size_t)w * 3 * h));
  rc = f->dec.enqueue(f->buf.p, f->stream);
  if (rc) return rc;
  FDL_CUDA_TRY(cudaStreamSynchronize(f->stream));
  rc = f->dec.check_status();
  if (rc) return rc;
  f->w = w; f->h = h;
  return FDL_OK;
} FDL_ABI_CATCH

int fdl_frame_image(const fdl_frame* f, fdl_image* out) try {
  if (!f || !out) return set_error(FDL_ERR_INVALID, "null argument");
  if (f->w <= 0) return set_error(FDL_ERR_INVALID, "nothing uploaded into this frame yet");
  out->data = f->buf.p; out->width = f->w; out->height = f->h; out->row_stride = (int64_t)f->w * 3; out->mem = FDL_MEM_DEVICE; out->_pad = 0;
  return FDL_OK;
} FDL_ABI_CATCH

}  // extern "C"
