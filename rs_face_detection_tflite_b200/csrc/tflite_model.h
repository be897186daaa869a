// tflite_model.h -- bounds-checked reader for the TFLite schema-v3 flatbuffers in models/.
//
// Replaces `FlatBufferModel::build_from_file` (face_detection.rs:188, face_landmark.rs:216,
// iris_landmark.rs:150): the host parses the .tflite file itself for weights and graph topology.
// Only the subset of the public schema (tensorflow/lite/schema/schema.fbs) the five dense graphs
// use is decoded.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fdl {

enum BuiltinOp : int {
  OP_ADD = 0, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4, OP_DEPTH_TO_SPACE = 5,
  OP_DEQUANTIZE = 6, OP_MAX_POOL_2D = 17, OP_RELU = 19, OP_RESHAPE = 22, OP_RESIZE_BILINEAR = 23,
  OP_PAD = 34, OP_PRELU = 54, OP_DENSIFY = 124
};
const char* op_name(int code);

enum TensorType : int { TT_F32 = 0, TT_F16 = 1, TT_I32 = 2, TT_U8 = 3, TT_I64 = 4 };

struct TfTensor {
  std::vector<int> shape;
  int type = 0;
  uint32_t buffer = 0;
  std::string name;
  // constant payload (view into the file image), empty for activations
  const uint8_t* data = nullptr;
  size_t nbytes = 0;
  bool has_sparsity = false;
  // SparsityParameters (Tensor field 6) of the sparse full-range detector's 1x1 weights: row-major CSR on the last
  // dimension -- `segments` over the flattened outer dimensions, `indices` = positions in the last dimension; `data`
  // then holds only the stored values.  sparse_ok is false for any other layout (-> the model is rejected).
  bool sparse_ok = false;
  std::vector<int32_t> sp_segments, sp_indices;
  int64_t elems() const { int64_t n = 1; for (int d : shape) n *= d; return n; }
};

struct TfOp {
  int code = -1;
  std::vector<int> inputs, outputs;
  // decoded builtin options (fields not applicable to `code` keep their defaults)
  int padding = 0;          // 0 SAME, 1 VALID
  int stride_w = 1, stride_h = 1;
  int dil_w = 1, dil_h = 1;
  int depth_multiplier = 1;
  int filter_w = 0, filter_h = 0;
  int fused_act = 0;
  int axis = 0;
  int block_size = 0;       // DEPTH_TO_SPACE
  std::vector<int> new_shape;
  bool align_corners = false, half_pixel_centers = false;
};

struct TfModel {
  uint32_t version = 0;
  std::vector<TfTensor> tensors;
  std::vector<TfOp> ops;
  std::vector<int> inputs, outputs;
  std::vector<uint8_t> file;  // owns the bytes the tensor payloads point into

  // Returns false and fills `err` on any malformed offset / unsupported construct.
  bool load(const std::string& path, std::string* err);
  // Constant tensor as f32 (widening f16 exactly, i.e. a folded DEQUANTIZE; sparse tensors are expanded first, i.e. a
  // folded DENSIFY: tensorflow/lite/kernels/densify.cc).
  bool const_f32(int tensor, std::vector<float>* out) const;
  bool const_i32(int tensor, std::vector<int>* out) const;
};

float half_to_float(uint16_t h);
// f32 -> f16, round to nearest even, saturating to the largest finite half (NaN stays NaN).
uint16_t float_to_half(float f);

}  // namespace fdl
