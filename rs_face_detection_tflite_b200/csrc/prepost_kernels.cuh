// prepost_kernels.cuh -- launch interface of the pre/post-processing kernels (sm_100a).
//
// They replace, on the device and batched over frames/faces/eyes:
//   * OpenCV getPerspectiveTransform / warpPerspective / copyMakeBorder / resize / flip and the
//     per-pixel normalisation loop of image_to_tensor (transform.rs:188-309),
//   * ssd_generate_anchors, decode_boxes, get_sigmoid_score, convert_to_detections
//     (face_detection.rs:269-413), weighted NMS (nms.rs:56-144), detection_letterbox_removal
//     (transform.rs:115-142),
//   * face_detection_to_roi (face_landmark.rs:180-198), the face-flag gate (:292-296),
//     project_landmarks (transform.rs:351-432), iris_roi_from_face_landmarks (iris_landmark.rs:268-292).
// This translation unit is compiled with -fmad=false (see glue_math.h).
#pragma once
#include <cuda_runtime.h>

#include "glue_math.h"

namespace fdl {

// ssd_generate_anchors -> out[n][2]
cudaError_t launch_anchors(const SsdOptions& opt, float* out, int n, cudaStream_t s);

// image_to_tensor setup for `n` slots.  rois == nullptr: full-frame ROI for every slot.
// slot_frame == nullptr: slot i reads frame i.  flip_mode: 0 never, 1 always, 2 odd slots (right eyes).
// slot_valid (optional): slots with 0 are marked invalid.  n_active (optional device counter).
cudaError_t launch_i2t_setup(const fdl_rect* rois, const int* slot_frame, const int* slot_valid, int n, int img_w, int img_h,
                             int out_w, int out_h, int keep_aspect, double range_min, double range_max, int flip_mode,
                             I2TParams* params, const int* n_active, cudaStream_t s);
// image_to_tensor pixels: frames = base of [F, img_h, row_stride] u8; out = [n, out_h, out_w, 3] f32
// (batch stride out_bstride floats); out_u8 optional [n, out_h, out_w, 3].
// compact / row_pos / compact_fstride (optional, rows_mode only): a device-resident copy of just the source rows the
// letterbox touches ([frame][compact row][row bytes], gathered by the copy engine); row_pos[src row] = compact row or -1.
// Source rectangle of a warp in frame pixels (inclusive; x1 < x0: empty).
// Zero-copy host frames: copy the (margin-grown) source rectangle of every face warp from the pinned host frames into the device
// frame buffer at the same offsets, and split the eye slots between the device copy and the host frames (see prepost_kernels.cu).
cudaError_t launch_roi_fill(const uint8_t* host_frames, uint8_t* dev_frames, long long frame_stride, long long row_stride, const I2TParams* params,
                            int n, const int* n_active, SrcBox* boxes, int frame_h, int max_ctas, cudaStream_t s,
                            const uint8_t* compact = nullptr, const int* row_pos = nullptr, long long compact_fstride = 0);
cudaError_t launch_eye_split(const I2TParams* eye_params, const SrcBox* face_boxes, int n, const int* n_active, I2TParams* p_dev, I2TParams* p_host,
                             cudaStream_t s);
cudaError_t launch_i2t(const uint8_t* frames, long long frame_stride, long long row_stride, const I2TParams* params, int n,
                       int out_w, int out_h, float* out, long long out_bstride, uint8_t* out_u8, const int* n_active,
                       cudaStream_t s, int rows_mode = 0, int src_w = 0, int max_ctas = 0,
                       const uint8_t* compact = nullptr, const int* row_pos = nullptr, long long compact_fstride = 0);
// rows_mode = 1 (detector letterbox): one CTA per output row, source rows staged in shared memory (see i2t_rows_kernel).

struct SsdPostArgs {
  const float* reg = nullptr; long long reg_bstride = 0;   // [B,N,16]
  const float* cls = nullptr; long long cls_bstride = 0;   // [B,N,1]
  const float* anchors = nullptr;                          // [N,2]
  int N = 0; int B = 0;
  float scale = 1.f;                                       // input height as f32 (face_detection.rs:259)
  const I2TParams* params = nullptr;                       // per-frame padding source (params[b].pad), or
  const double* padding4 = nullptr;                        // explicit per-frame padding [B,4]
  // outputs, addressed with byte strides so that they can live inside fdl_frame_result records or in flat arrays:
  char* det_base = nullptr; long long det_stride = 0;      // frame b: fdl_detection[max_out] at det_base + b*det_stride
  char* ndet_base = nullptr; long long ndet_stride = 0;    // frame b: int32 count (clamped to max_out) at ndet_base + b*ndet_stride
  int max_out = FDL_MAX_DETECTIONS;
  int* n_total = nullptr; long long n_total_stride = 1;    // optional: frame b's number of clusters before the max_out cap at n_total[b*n_total_stride]
  int32_t* surv_anchor = nullptr; int32_t* surv_cluster = nullptr; int cap_surv = 0; int* n_surv = nullptr;  // optional debug outputs
};
cudaError_t launch_ssd_postprocess(const SsdPostArgs& a, cudaStream_t s);

// Compacts the first min(n_detections, max_faces) detections of every frame into face slots
// (frame-major order): slot_frame/slot_face [B*max_faces], *n_faces, *n_eyes = 2 * *n_faces.
cudaError_t launch_face_select(fdl_frame_result* frames, int B, int max_faces, int* slot_frame, int* slot_face, int* n_faces,
                               int* n_eyes, cudaStream_t s);
// face_detection_to_roi per face slot -> rois[slot], also stored into face_results[frame*max_faces+face].face_roi.
cudaError_t launch_face_roi(const fdl_frame_result* frames, const int* slot_frame, const int* slot_face, int max_slots, int max_faces,
                            int img_w, int img_h, fdl_rect* rois, int* slot_valid, fdl_face_result* faces, const int* n_faces,
                            cudaStream_t s);
// Landmark post-processing per face slot: face-flag gate, project 468 landmarks, eye ROIs
// (eye slot 2*slot = left, 2*slot+1 = right).
cudaError_t launch_landmark_post(const float* raw, long long raw_bstride, const float* flag, long long flag_bstride,
                                 const I2TParams* params, const fdl_rect* rois, const int* slot_frame, const int* slot_face,
                                 int max_slots, int max_faces, int tensor_w, int tensor_h, fdl_face_result* faces, fdl_rect* eye_rois,
                                 int* eye_frame, int* eye_valid, const int* n_faces, cudaStream_t s, int refine = 0);
// Iris post-processing per eye slot: project 71 + 5 landmarks (flip for right eyes).
cudaError_t launch_iris_post(const float* contour, long long contour_bstride, const float* iris, long long iris_bstride,
                             const I2TParams* params, const fdl_rect* eye_rois, const int* eye_valid, const int* slot_frame,
                             const int* slot_face, int max_eye_slots, int max_faces, int tensor_w, int tensor_h,
                             fdl_face_result* faces, const int* n_eyes, cudaStream_t s, int refine = 0, double focal_length_mm = 0.0);
// Stand-alone iris refinement helpers (iris_landmark.rs:380-433) on f64 landmark triples.
cudaError_t launch_refine_landmarks(const double* face, const double* left, int n_left, const double* right, int n_right, double* out,
                                    cudaStream_t s);
cudaError_t launch_iris_metrics(const double* iris, int img_w, int img_h, double focal_length_mm, double iris_size_px, double* out2,
                                cudaStream_t s);

// Stand-alone helpers behind the free functions of the C ABI (one thread each).
cudaError_t launch_face_detection_to_roi(const fdl_detection* det, int img_w, int img_h, int size_mode, fdl_rect* out, int* ok,
                                         cudaStream_t s);
cudaError_t launch_eye_rois(const double* lm4xy /*[8]: 33,133,362,263 (x,y)*/, int img_w, int img_h, fdl_rect* out2, int* ok,
                            cudaStream_t s);
cudaError_t launch_project(const float* raw, int n, int tensor_w, int tensor_h, int img_w, int img_h, const double* pad4,
                           const fdl_rect* roi_or_null, int flip, float* out, cudaStream_t s);

}  // namespace fdl
