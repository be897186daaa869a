// plan.h -- turns a parsed .tflite graph into a list of fused kernel launches ("steps").
//
// This is what replaces TFLite's InterpreterBuilder/allocate_tensors (face_detection.rs:207-210,
// face_landmark.rs:233-236, iris_landmark.rs:161-164): instead of interpreting the 97..354 ops one
// by one, the planner
//   * folds DEQUANTIZE (f16 -> f32 constants) and DENSIFY (CSR -> dense constants) at load,
//   * folds a spatial zero PAD into the VALID depthwise / convolution that consumes it (explicit padding) and a fused
//     RELU of CONV_2D / ADD into the step's activation (both forms occur in the sparse full-range detector),
//   * turns RESHAPE / CONCATENATION(axis=1) into aliases so heads write straight into the
//     [B,N,16] / [B,N,1] outputs,
//   * pattern-matches  DW3x3 -> CONV1x1 -> [ADD skip] -> [RELU|PRELU]  (BlazeBlock) with
//     the skip branch's MAX_POOL / channel PAD folded into the epilogue,
//   * fuses CONV -> [RELU|PRELU] and RESIZE_BILINEAR -> ADD,
//   * assigns activation buffers by liveness (per-batch-item offsets, scaled by B at run time).
// The plan is device-independent (pure host data) so it can be inspected and tested without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "chain.h"
#include "tflite_model.h"

namespace fdl {

enum StepKind : int {
  STEP_CONV = 0,     // CONV_2D (any kh,kw,stride) + bias [+ act]                     -> fused_conv kernel
  STEP_BLOCK = 1,    // DW3x3(stride) + bias -> CONV1x1 + bias [+ skip] [+ act]       -> fused_conv kernel (DW prologue)
  STEP_DW = 2,       // standalone DEPTHWISE_CONV_2D [+ act]
  STEP_POOL = 3,     // standalone MAX_POOL_2D
  STEP_PADC = 4,     // standalone channel PAD
  STEP_ADD = 5,      // standalone ADD [+ act]
  STEP_ACT = 6,      // standalone RELU / PRELU
  STEP_RESIZE = 7,   // RESIZE_BILINEAR (half pixel centres) [+ ADD other] [+ act]
  STEP_D2S = 8       // DEPTH_TO_SPACE (block size in `stride`): the sparse full-range detector's heads
};
enum ActKind : int { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2 };

// A view of a [B,H,W,C] f32 tensor inside the activation arena.  Element (b,y,x,c) lives at
//   arena + buf_offset*B + b*batch_stride + offset + (y*W + x)*C + c
// (buf_offset, batch_stride, offset in floats; buf_offset is per batch item and scaled by B).
struct TensorRef {
  int tensor = -1;           // tflite tensor index (for diagnostics)
  int64_t buf_offset = 0;    // start of the owning buffer, per batch item
  int64_t batch_stride = 0;  // floats between consecutive batch items in the owning buffer
  int64_t offset = 0;        // offset inside one batch item of the owning buffer
  int H = 0, W = 0, C = 0;
};

struct Step {
  int kind = STEP_CONV;
  TensorRef in, out;
  TensorRef skip;            // BLOCK: residual source; ADD/RESIZE: second operand (tensor == -1: none)
  int skip_pool = 0;         // BLOCK: residual = MAX_POOL 2x2 s2 of skip
  int skip_c = 0;            // BLOCK: channels of the residual source (< out.C: zero channel PAD)
  int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;   // CONV / DW / POOL geometry (DW part of BLOCK)
  int act = ACT_NONE;
  // offsets (in floats) into the weight arena; -1 = absent
  int64_t w_dw = -1, b_dw = -1;   // DW weights [kh*kw][C], bias [C]
  int64_t w = -1, b = -1;         // CONV/PW weights [K4][Npad] (K = kh*kw*Cin, Npad = roundup(Cout,4)), bias [Npad]
  int64_t alpha = -1;             // PRELU slopes [C]
  // BLOCK steps also carry the pointwise weights packed for tcgen05 (UMMA K-major core matrices):
  // [wsplit][Cin/4][Np][4] floats; wsplit == 1 when every weight is tf32-exact (f16-stored detectors),
  // else 2 = (tf32(w), w - tf32(w)).  -1 when Cin % 8 != 0.
  int64_t w_umma = -1;
  int Np = 0, wsplit = 1;
  // The same pointwise weights for tcgen05 kind::f16 with the A operand split as (f16 hi, f16 lo): K' = 2*Cin, plane q holds,
  // for output channel n, the 8 halves (w[4q..4q+3], w[4q..4q+3]) -- the same weight against the hi and the lo half of the
  // activation.  [wsplit16][Cin/4][Np][8 halves]; wsplit16 == 1 when every weight is f16-exact (the f16-stored detectors),
  // else 2: second copy = (w - f16(w) as f16, 0).  Stored as raw bits, two halves per float.  -1 when Cin % 8 != 0.
  int64_t w_f16 = -1;
  int wsplit16 = 1;
  // CONV and BLOCK steps with Cin % 4 == 0: weights for the general tensor-core kernel,
  // [n_tiles][wsplit][Kp/4][Nt][4] with Kp = roundup(K,32), Nt = min(roundup(N,16),128).
  int64_t w_tc = -1;
  int Kp = 0, Nt = 0, n_tiles = 0;
  int K = 0, K4 = 0, N = 0, Npad = 0;  // contraction size (K4 = roundup(K,4) rows stored) / output channels of the CONV/PW part
  // Branch of the step: 0 = trunk (the step feeds several graph outputs) or the branch of graph output 0; k >= 1 = the step
  // feeds graph output k only.  Steps with stream >= 1 may run on an auxiliary CUDA stream beside the others (Net::forward):
  // the planner never recycles a buffer such a step reads or writes, so stream order is the only ordering they need.
  int stream = 0;
  std::vector<int> ops;           // tflite op indices folded into this step
  std::string text;               // human-readable description
};

struct Plan {
  std::vector<Step> steps;
  std::vector<float> weights;     // host copy of the weight arena
  TensorRef input;
  std::vector<TensorRef> outputs; // graph output order == interpreter.outputs() order
  int64_t arena_per_item = 0;     // floats of activation arena per batch item
  int64_t algo_bytes_per_item = 0;  // sum over steps of (input + skip + output) bytes: the block-fused floor
  int64_t flops_per_item = 0;
  int num_tflite_ops = 0;
  int num_streams = 1;            // 1 + number of auxiliary branches (see Step::stream)
  std::vector<ChainPlan> chains;  // the steps on maps of at most 8 x 8 pixels as one launch per chain (chain.h): none, one or two

  bool build(const TfModel& m, std::string* err);
  std::string describe() const;
};

}  // namespace fdl
