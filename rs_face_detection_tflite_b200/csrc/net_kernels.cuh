// net_kernels.cuh -- launch interface of the network kernels (sm_100a).
//
// These kernels replace the TFLite CPU kernels behind `interpreter.invoke()`
// (face_detection.rs:235, face_landmark.rs:265, iris_landmark.rs:203); op semantics follow
// SURVEY.md Appendix A.3.  All tensors are NHWC f32, batched over frames/faces/eyes.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace fdl {

struct TView {           // a [B,H,W,C] f32 view
  float* p = nullptr;    // element (0,0,0,0)
  long long bstride = 0; // floats between batch items
  int H = 0, W = 0, C = 0;
};

struct ConvArgs {
  TView in, out, skip;
  int mode = 0;          // 0: CONV_2D (im2col on the fly); 1: BLOCK (depthwise 3x3 prologue + pointwise)
  int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0;
  int K = 0, K4 = 0, N = 0, Npad = 0;
  const float* w = nullptr;     // [K4][Npad]
  const float* bias = nullptr;  // [Npad]
  const float* w_dw = nullptr;  // [9][Cin]
  const float* b_dw = nullptr;  // [Cin]
  const float* alpha = nullptr; // [N] (PRELU)
  int act = 0;
  int has_skip = 0, skip_pool = 0, skip_c = 0;
  int B = 0;
  const int* n_active = nullptr;  // optional device counter: only the first *n_active batch items are computed
  int mma = 0;                    // pointwise arithmetic: 0 fp32 FFMA, 1 split-TF32 tensor cores
};

struct EltArgs {
  TView in, out, other;
  int kind = 0;          // StepKind
  int stride = 1, pad_t = 0, pad_l = 0;
  const float* w_dw = nullptr;
  const float* b_dw = nullptr;
  const float* alpha = nullptr;
  int act = 0;
  int has_other = 0;
  int B = 0;
  const int* n_active = nullptr;
};

// Streaming fp32 pointwise convolution for the large maps (pw_kernel.cu).
struct Step;
cudaError_t pw_stream_init();
bool pw_stream_supported(const Step& s, int B);
cudaError_t launch_pw_stream(const ConvArgs& a, cudaStream_t stream);

// Returns cudaSuccess or the launch error.  `stream` is the handle's stream.
cudaError_t launch_fused_conv(const ConvArgs& a, cudaStream_t stream);
cudaError_t launch_elementwise(const EltArgs& a, cudaStream_t stream);
// RGB stem convolutions (stem_kernel.cu)
bool stem_supported(const ConvArgs& a);
cudaError_t launch_stem_conv(const ConvArgs& a, cudaStream_t stream);
cudaError_t stem_kernels_init();
// the same on the tensor cores (stem_tc_kernel.cu): Cout <= 32, output size a multiple of 8 x 16, ConvArgs::mma set
bool stem_tc_supported(const ConvArgs& a);
cudaError_t launch_stem_tc(const ConvArgs& a, cudaStream_t stream);
cudaError_t stem_tc_init();
// large-map 1x1 convolutions as a streaming tensor-core GEMM (pw_tc_kernel.cu): Cin 64 / 128, Cout <= 64, dense tensors
struct Step;
bool pw_tc_supported(const Step& s, int B);
cudaError_t launch_pw_tc(const ConvArgs& a, cudaStream_t stream);
cudaError_t pw_tc_init();
cudaError_t net_kernels_init();   // opt-in shared memory sizes; call once per device

void count_launch();              // bumps the library-wide launch counter (fdl_launch_count)

}  // namespace fdl
