// net.h -- a planned .tflite graph resident on one B200: weights uploaded once, activation arena
// sized for the largest batch seen, one launch per fused step.  Replaces the per-call
// InterpreterBuilder / allocate_tensors / invoke sequence of the reference
// (face_detection.rs:207-235, face_landmark.rs:233-265, iris_landmark.rs:161-203).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "net_kernels.cuh"
#include "plan.h"

namespace fdl {

class Net {
 public:
  // device < 0: plan-only handle (describe / introspection, cannot run).
  static Net* create(const std::string& path, int device, std::string* err, int* code);
  ~Net();

  const Plan& plan() const { return plan_; }
  int device() const { return device_; }
  int64_t in_elems() const { return (int64_t)plan_.input.H * plan_.input.W * plan_.input.C; }
  int64_t out_elems(int i) const { const TensorRef& r = plan_.outputs[i]; return (int64_t)r.H * r.W * r.C; }
  int num_outputs() const { return (int)plan_.outputs.size(); }
  void set_mode(int m) { mode_ = m; }
  int mode() const { return mode_; }

  // Grow the activation arena so that batches up to B fit.  Invalidates earlier views.
  bool reserve(int B, std::string* err);
  // Views for batch size B (buffers are [B, item] contiguous; offsets scale with B).
  TView view(const TensorRef& r, int B) const;
  TView input_view(int B) const { return view(plan_.input, B); }
  TView output_view(int i, int B) const { return view(plan_.outputs[i], B); }

  // Enqueue every step for batch B on `stream`.  n_active (optional, device pointer): only the
  // first *n_active items are computed (data-dependent fan-out without a host round trip).
  // input_override (optional): read the network input from this [B, in_elems] buffer instead of the arena.
  // step_events (optional): num_steps + 1 events, recorded before every step and after the last one.
  cudaError_t forward(int B, cudaStream_t stream, const int* n_active = nullptr, const float* input_override = nullptr,
                      cudaEvent_t* step_events = nullptr);

 private:
  Net() = default;
  Plan plan_;
  int device_ = -1;
  int mode_ = 1;   // 1: tensor-core BlazeBlock kernel where it applies (default), 0: fp32 FFMA kernels only
  float* d_weights_ = nullptr;
  float* d_arena_ = nullptr;
  int cap_B_ = 0;
  // Branch streams (Step::stream >= 1): the steps that feed only graph output k run on aux_[k-1], forked from / joined
  // into the caller's stream with events, so the two heads of the landmark / iris / detector graphs run side by side.
  std::vector<cudaStream_t> aux_;
  std::vector<cudaEvent_t> ev_fork_, ev_join_;

};

}  // namespace fdl
