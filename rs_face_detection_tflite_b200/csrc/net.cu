// net.cu -- see net.h.
#include "net.h"

#include <cstdlib>
#include <utility>

#include "fdl_status.h"
#include "chain_kernel.cuh"
#include "mma_kernels.cuh"

namespace fdl {

Net* Net::create(const std::string& path, int device, std::string* err, int* code) {
  TfModel m;
  if (!m.load(path, err)) { *code = err->find("cannot open") != std::string::npos ? FDL_ERR_IO : FDL_ERR_MODEL; return nullptr; }
  Net* n = new Net();
  if (!n->plan_.build(m, err)) { *code = FDL_ERR_MODEL; delete n; return nullptr; }
  n->device_ = device;
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = net_kernels_init();
    if (e == cudaSuccess) e = mma_kernels_init();
    if (e == cudaSuccess) e = conv_tc_init();
    if (e == cudaSuccess) e = block_ws_init();
    if (e == cudaSuccess) e = pw_stream_init();
    if (e == cudaSuccess) e = chain_init();
    if (e == cudaSuccess) e = cudaMalloc(&n->d_weights_, n->plan_.weights.size() * sizeof(float));
    if (e == cudaSuccess)
      e = cudaMemcpy(n->d_weights_, n->plan_.weights.data(), n->plan_.weights.size() * sizeof(float), cudaMemcpyHostToDevice);
    static const bool branch_streams = getenv("FDL_BRANCH_STREAMS") ? atoi(getenv("FDL_BRANCH_STREAMS")) != 0 : true;
    for (int k = 1; branch_streams && k < n->plan_.num_streams && e == cudaSuccess; ++k) {
      cudaStream_t st = nullptr; cudaEvent_t f = nullptr, j = nullptr;
      e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j, cudaEventDisableTiming);
      n->aux_.push_back(st); n->ev_fork_.push_back(f); n->ev_join_.push_back(j);
    }
    if (e != cudaSuccess) {
      *err = std::string("CUDA: ") + cudaGetErrorString(e);
      *code = FDL_ERR_CUDA;
      delete n;
      return nullptr;
    }
  }
  return n;
}

Net::~Net() {
  if (device_ >= 0) {
    cudaSetDevice(device_);
    if (d_weights_) cudaFree(d_weights_);
    if (d_arena_) cudaFree(d_arena_);
    for (cudaEvent_t e : ev_fork_) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_join_) if (e) cudaEventDestroy(e);
    for (cudaStream_t st : aux_) if (st) cudaStreamDestroy(st);
  }
}

bool Net::reserve(int B, std::string* err) {
  if (device_ < 0) { *err = "plan-only handle cannot run (no CUDA device bound)"; return false; }
  if (B <= cap_B_) return true;
  cudaSetDevice(device_);
  if (d_arena_) { cudaDeviceSynchronize(); cudaFree(d_arena_); d_arena_ = nullptr; cap_B_ = 0; }
  cudaError_t e = cudaMalloc(&d_arena_, (size_t)plan_.arena_per_item * (size_t)B * sizeof(float));
  if (e != cudaSuccess) { *err = std::string("CUDA: arena allocation failed: ") + cudaGetErrorString(e); return false; }
  cap_B_ = B;
  return true;
}

TView Net::view(const TensorRef& r, int B) const {
  TView v;
  v.p = d_arena_ + r.buf_offset * (int64_t)B + r.offset;
  v.bstride = r.batch_stride;
  v.H = r.H; v.W = r.W; v.C = r.C;
  return v;
}

cudaError_t Net::forward(int B, cudaStream_t stream, const int* n_active, const float* input_override, cudaEvent_t* step_events) {
  auto in_view = [&](const Step& st) {
    TView v = view(st.in, B);
    if (input_override && st.in.tensor == plan_.input.tensor) { v.p = const_cast<float*>(input_override); v.bstride = in_elems(); }
    return v;
  };
  if (B > cap_B_ || device_ < 0) return cudaErrorInvalidValue;
  size_t step_index = 0;
  // Branch streams: per-launch timing (step_events) keeps everything on the caller's stream.
  cudaStream_t const main_stream = stream;
  const bool multi = !aux_.empty() && !step_events;
  std::vector<int> synced(aux_.size() + 1, -2);      // last step of the main stream each auxiliary stream has waited for
  std::vector<char> used(aux_.size() + 1, 0);
  std::vector<std::pair<int64_t, int>> main_writes;  // (root buffer, step) written on the main stream so far
  int main_last = -1;                                // last step enqueued on the main stream
  int si = -1;
  const bool chained = mode_ == 1 && !plan_.chains.empty() && chain_enabled();
  for (const Step& s : plan_.steps) {
    cudaError_t e;
    ++si;
    stream = main_stream;
    const ChainPlan* ch = nullptr;
    if (chained)
      for (const ChainPlan& c : plan_.chains) if (si >= c.first_step && si <= c.last_step) ch = &c;
    if (ch) {
      // a tail chain: ONE launch on the caller's stream at the position of its first step (everything before it is on that
      // stream too); the steps after it see its stores as main-stream writes of its last step
      if (step_events && (e = cudaEventRecord(step_events[step_index++], stream)) != cudaSuccess) return e;
      if (multi) { main_writes.push_back({s.out.buf_offset, ch->last_step}); main_last = ch->last_step; }
      if (si != ch->first_step) continue;
      ChainArgs a;
      for (size_t k = 0; k < ch->ops.size(); ++k) a.ops[k] = ch->ops[k];
      for (size_t k = 0; k < ch->loads.size(); ++k) a.loads[k] = ch->loads[k];
      a.n_ops = (int)ch->ops.size(); a.n_loads = (int)ch->loads.size();
      a.weights = d_weights_; a.arena = d_arena_; a.B = B; a.items = ch->items; a.n_active = n_active;
      if ((e = launch_chain(a, stream)) != cudaSuccess) return e;
      continue;
    }
    if (multi && s.stream >= 1 && s.stream <= (int)aux_.size()) {
      const int k = s.stream;
      stream = aux_[k - 1];
      // read-after-write across streams: the newest main-stream writer of the buffers this step reads (the network input,
      // *n_active and anything else produced before this pass count as "before step 0")
      int need = -1;
      for (const TensorRef* r : {&s.in, &s.skip}) {
        if (r->tensor < 0) continue;
        for (const auto& w : main_writes) if (w.first == r->buf_offset && w.second > need) need = w.second;
      }
      if (need > synced[k]) {
        if ((e = cudaEventRecord(ev_fork_[k - 1], main_stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(stream, ev_fork_[k - 1], 0)) != cudaSuccess) return e;
        synced[k] = main_last;
      }
      used[k] = 1;
    } else if (multi) {
      main_writes.push_back({s.out.buf_offset, si});
      main_last = si;
    }
    if (step_events && (e = cudaEventRecord(step_events[step_index++], stream)) != cudaSuccess) return e;
    if (mode_ >= 1 && block_tc_supported(s)) {
      BlockTcLaunch l;
      TView in = in_view(s), out = view(s.out, B);
      l.in = in.p; l.out = out.p;
      BlockTcArgs& a = l.args;
      a.w_umma = d_weights_ + s.w_umma; a.bias = d_weights_ + s.b; a.w_dw = d_weights_ + s.w_dw; a.b_dw = d_weights_ + s.b_dw;
      a.alpha = s.alpha >= 0 ? d_weights_ + s.alpha : nullptr;
      l.alpha_host = s.alpha >= 0 ? plan_.weights.data() + s.alpha : nullptr;
      l.bias_host = plan_.weights.data() + s.b;
      if (s.w_f16 >= 0) { a.w_f16 = d_weights_ + s.w_f16; a.wsplit16 = s.wsplit16; }
      a.C = s.in.C; a.N = s.out.C; a.Np = s.Np; a.H = s.out.H; a.W = s.out.W; a.B = B;
      a.act = s.act; a.stride = s.stride; a.wsplit = s.wsplit; a.n_active = n_active;
      if (s.skip.tensor >= 0) {
        TView sk = view(s.skip, B);
        a.skip_c = s.skip_c;
        if (s.skip_pool && s.skip.tensor == s.in.tensor && s.stride == 2) a.skip_mode = 4;
        else if (s.skip_pool) { a.skip_mode = 3; a.skip = sk.p; a.skip_bstride = sk.bstride; }
        else if (s.skip.tensor == s.in.tensor && s.stride == 1) a.skip_mode = 1;
        else { a.skip_mode = 2; a.skip = sk.p; a.skip_bstride = sk.bstride; }
      }
      e = (mode_ == 1 && block_ws_supported(s)) ? launch_block_ws(l, stream) : launch_block_tc(l, stream);
    } else if (mode_ >= 1 && pw_tc_supported(s, B)) {
      ConvArgs a;
      a.in = in_view(s); a.out = view(s.out, B);
      a.kh = s.kh; a.kw = s.kw; a.stride = s.stride; a.K = s.K; a.K4 = s.K4; a.N = s.N; a.Npad = s.Npad;
      a.w = d_weights_ + s.w; a.bias = d_weights_ + s.b;
      if (s.alpha >= 0) a.alpha = d_weights_ + s.alpha;
      a.act = s.act; a.B = B; a.n_active = n_active;
      e = launch_pw_tc(a, stream);
    } else if (mode_ >= 1 && pw_stream_supported(s, B)) {
      ConvArgs a;
      a.in = in_view(s); a.out = view(s.out, B);
      a.kh = a.kw = 1; a.K = s.K; a.K4 = s.K4; a.N = s.N; a.Npad = s.Npad;
      a.w = d_weights_ + s.w; a.bias = d_weights_ + s.b;
      if (s.alpha >= 0) a.alpha = d_weights_ + s.alpha;
      a.act = s.act; a.B = B; a.n_active = n_active;
      e = launch_pw_stream(a, stream);
    } else if (mode_ >= 1 && conv_tc_supported(s)) {
      ConvTcArgs a;
      TView in = in_view(s), out = view(s.out, B);
      a.in.p = in.p; a.in.bstride = in.bstride; a.in.H = in.H; a.in.W = in.W; a.in.C = in.C;
      a.out.p = out.p; a.out.bstride = out.bstride; a.out.H = out.H; a.out.W = out.W; a.out.C = out.C;
      a.mode = s.kind == STEP_BLOCK ? 1 : (s.in.C % 4 == 0 ? 0 : 2);
      a.kh = s.kh; a.kw = s.kw; a.stride = s.stride; a.pad_t = s.pad_t; a.pad_l = s.pad_l;
      a.K = s.K; a.Kp = s.Kp; a.N = s.N; a.Nt = s.Nt; a.n_tiles = s.n_tiles;
      a.w_tc = d_weights_ + s.w_tc; a.bias = d_weights_ + s.b;
      if (s.w_dw >= 0) { a.w_dw = d_weights_ + s.w_dw; a.b_dw = d_weights_ + s.b_dw; }
      if (s.alpha >= 0) a.alpha = d_weights_ + s.alpha;
      a.act = s.act; a.wsplit = s.wsplit;
      if (s.skip.tensor >= 0) {
        TView sk = view(s.skip, B);
        a.has_skip = 1; a.skip.p = sk.p; a.skip.bstride = sk.bstride; a.skip.H = sk.H; a.skip.W = sk.W; a.skip.C = sk.C;
        a.skip_pool = s.skip_pool; a.skip_c = s.skip_c;
      }
      a.B = B; a.n_active = n_active;
      e = launch_conv_tc(a, stream);
    } else if (s.kind == STEP_CONV || s.kind == STEP_BLOCK) {
      ConvArgs a;
      a.in = in_view(s); a.out = view(s.out, B);
      a.mode = s.kind == STEP_BLOCK ? 1 : 0;
      a.kh = s.kh; a.kw = s.kw; a.stride = s.stride; a.pad_t = s.pad_t; a.pad_l = s.pad_l;
      a.K = s.K; a.K4 = s.K4; a.N = s.N; a.Npad = s.Npad;
      a.w = d_weights_ + s.w; a.bias = d_weights_ + s.b;
      if (s.w_dw >= 0) { a.w_dw = d_weights_ + s.w_dw; a.b_dw = d_weights_ + s.b_dw; }
      if (s.alpha >= 0) a.alpha = d_weights_ + s.alpha;
      a.act = s.act;
      if (s.skip.tensor >= 0) { a.has_skip = 1; a.skip = view(s.skip, B); a.skip_pool = s.skip_pool; a.skip_c = s.skip_c; }
      a.B = B; a.n_active = n_active; a.mma = mode_ >= 1 ? 1 : 0;
      e = launch_fused_conv(a, stream);
    } else {
      EltArgs a;
      a.in = in_view(s); a.out = view(s.out, B);
      a.kind = s.kind; a.stride = s.stride; a.pad_t = s.pad_t; a.pad_l = s.pad_l;
      if (s.w_dw >= 0) { a.w_dw = d_weights_ + s.w_dw; a.b_dw = d_weights_ + s.b_dw; }
      if (s.alpha >= 0) a.alpha = d_weights_ + s.alpha;
      a.act = s.act;
      if (s.skip.tensor >= 0) { a.has_other = 1; a.other = view(s.skip, B); }
      a.B = B; a.n_active = n_active;
      e = launch_elementwise(a, stream);
    }
    if (e != cudaSuccess) return e;
  }
  stream = main_stream;
  for (size_t k = 1; k < used.size(); ++k) {
    if (!used[k]) continue;
    cudaError_t e = cudaEventRecord(ev_join_[k - 1], aux_[k - 1]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, ev_join_[k - 1], 0);
    if (e != cudaSuccess) return e;
  }
  if (step_events) return cudaEventRecord(step_events[step_index], stream);
  return cudaSuccess;
}

}  // namespace fdl
