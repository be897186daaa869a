/*
 * fdl.h -- C ABI of the B200-native detect -> landmark -> iris path.
 *
 * This is the drop-in boundary for rs-face-detection-tflite's inference path: every entry point
 * below replaces one call a reference user makes into `face_detection_lite::*` (which in the
 * reference fans out into the TFLite C++ interpreter and OpenCV).  Below this header: C++ host
 * code + hand-written sm_100a CUDA kernels (no TFLite, no OpenCV, no CPU fallback -- a call made
 * without a usable CUDA device returns FDL_ERR_CUDA).  Above it: the Rust shim crate
 * (rust_shim/, source only), the C++ mirror (include/fdl.hpp) and the Python ctypes mirror
 * (rs_face_detection_tflite_b200/api.py).
 *
 * Reference citations are relative to /root/reference/src/face_detection_lite/.
 *
 * Conventions
 *   - every function returning `int` returns FDL_OK (0) or a negative FDL_ERR_* code; the message
 *     is available through fdl_last_error() (thread-local).  Nothing aborts, nothing throws
 *     across the ABI (the reference mixes anyhow::Error and panics).
 *   - images are 8UC3 **RGB**, HWC, row-major, like the `Mat` produced by
 *     utils.rs:8-21 `convert_image_to_mat`; `row_stride` is in bytes.
 *   - a handle is bound to one CUDA device and one stream; calls on one handle must be
 *     serialised by the caller, different handles are independent.
 */
#ifndef FDL_H_
#define FDL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FDL_API
#else
#define FDL_API __attribute__((visibility("default")))
#endif

/* ---------------------------------------------------------------- status codes */
enum {
  FDL_OK = 0,
  FDL_ERR_INVALID = -1,      /* bad argument (null pointer, capacity, "bbox must be normalized" transform.rs:52) */
  FDL_ERR_IO = -2,           /* model file missing / unreadable (face_detection.rs:188) */
  FDL_ERR_MODEL = -3,        /* malformed or unsupported .tflite ("unsupported model type" face_detection.rs:184,
                                "incompatible model" face_landmark.rs:246, iris_landmark.rs:172-184) */
  FDL_ERR_CUDA = -4,         /* no device / CUDA runtime error: the product never falls back to the CPU */
  FDL_ERR_CAPACITY = -5,     /* caller-provided output buffer too small */
  FDL_ERR_INTERNAL = -6
};

/* ---------------------------------------------------------------- value types (types.rs) */

/* types.rs:24-37 `Rect` (f64 fields; rotation in radians, clockwise). */
typedef struct fdl_rect {
  double x_center, y_center, width, height, rotation;
  int32_t normalized;
  int32_t _pad;
} fdl_rect;

/* types.rs:189-246 `Detection`: data = Array2<f32>[8,2] row-major
 * (row0 = xmin,ymin; row1 = xmax,ymax; rows 2..7 = keypoints, face_detection.rs:91-98) + score.
 * `anchor` is not in the reference: the SSD anchor index of the NMS cluster's top detection
 * ("kept index"). */
typedef struct fdl_detection {
  float data[16];
  float score;
  int32_t anchor;
} fdl_detection;

/* types.rs:176-187 `Landmark` (f64). */
typedef struct fdl_landmark {
  double x, y, z;
} fdl_landmark;

/* The `&Mat` argument of every infer(): 8UC3 RGB. */
enum { FDL_MEM_HOST = 0, FDL_MEM_DEVICE = 1 };
typedef struct fdl_image {
  const uint8_t* data;
  int32_t width, height;
  int64_t row_stride;   /* bytes; 0 means width*3 */
  int32_t mem;          /* FDL_MEM_HOST (pageable or pinned) or FDL_MEM_DEVICE (on the handle's device) */
  int32_t _pad;
} fdl_image;

/* face_detection.rs:117-123 `FaceDetectionModel`.  FullSparse (4): the planner folds DENSIFY (CSR
 * f16 weights) and the spatial PADs at load and runs DEPTH_TO_SPACE on the device. */
enum {
  FDL_MODEL_FRONT_CAMERA = 0,
  FDL_MODEL_BACK_CAMERA = 1,
  FDL_MODEL_SHORT = 2,
  FDL_MODEL_FULL = 3,
  FDL_MODEL_FULL_SPARSE = 4
};

/* transform.rs:15-24 `SizeMode`; FDL_SIZE_MODE_NONE mirrors `Option::None` (-> SquareLong,
 * face_landmark.rs:188). */
enum { FDL_SIZE_MODE_NONE = -1, FDL_SIZE_MODE_DEFAULT = 0, FDL_SIZE_MODE_SQUARE_LONG = 1, FDL_SIZE_MODE_SQUARE_SHORT = 2 };

enum {
  FDL_NUM_FACE_LANDMARKS = 468,     /* face_landmark.rs:27 */
  FDL_NUM_EYE_CONTOUR = 71,         /* iris_landmark.rs:206-228: 213 / 3 */
  FDL_NUM_IRIS = 5,                 /* 15 / 3 */
  FDL_MAX_DETECTIONS = 32           /* capacity of the on-device NMS output per frame */
};

typedef struct fdl_detector fdl_detector;
typedef struct fdl_landmark_model fdl_landmark_model;
typedef struct fdl_iris_model fdl_iris_model;
typedef struct fdl_net fdl_net;
typedef struct fdl_pipeline fdl_pipeline;
typedef struct fdl_jpeg_decoder fdl_jpeg_decoder;
typedef struct fdl_pool fdl_pool;
typedef struct fdl_frame fdl_frame;

/* ---------------------------------------------------------------- library */
FDL_API const char* fdl_last_error(void);
FDL_API const char* fdl_version(void);
/* Number of CUDA devices visible (0 when there is none; never an error). */
FDL_API int fdl_device_count(void);
/* Number of kernel launches issued by this library on the calling thread's handles since load
 * (monotonic; used by bench.py for "gpu_launches"). */
FDL_API uint64_t fdl_launch_count(void);

/* ---------------------------------------------------------------- FaceDetection */
/* FaceDetection::new(model_type, model_path) face_detection.rs:153-195.  `model_dir` is a
 * DIRECTORY (NULL -> "./models") to which the per-variant file name is appended, as in the
 * reference.  Parses the .tflite flatbuffer, plans fused kernels, uploads weights, builds the
 * SSD anchor table (ssd_generate_anchors, :366-413) on the device. */
FDL_API int fdl_detector_create(int model, const char* model_dir, int device, fdl_detector** out);
FDL_API void fdl_detector_destroy(fdl_detector*);
FDL_API int fdl_detector_input_size(const fdl_detector*);   /* S: 128 / 256 / 192 */
FDL_API int fdl_detector_num_anchors(const fdl_detector*);  /* N: 896 / 2304 */
/* Copy the [N,2] f32 anchor table (x,y centres) to `out_xy` (host). */
FDL_API int fdl_detector_anchors(const fdl_detector*, float* out_xy, int cap_anchors);
/* FaceDetection::infer(&self, &Mat, Option<Rect>) face_detection.rs:205-267:
 * image_to_tensor (letterbox, range -1..1) -> network -> decode_boxes -> sigmoid -> threshold ->
 * weighted NMS -> letterbox removal, all on the device.  `roi` may be NULL. */
FDL_API int fdl_detector_infer(fdl_detector*, const fdl_image* image, const fdl_rect* roi,
                               fdl_detection* out, int cap, int* n_out);
/* Batched variant: `batch` images of identical size; out[i*cap .. i*cap+n_out[i]). */
FDL_API int fdl_detector_infer_batch(fdl_detector*, const fdl_image* images, int batch,
                                     fdl_detection* out, int cap_per_image, int* n_out);

/* Stage-level entry points (parity protocol steps 2 and 3). */
/* interpreter.invoke() face_detection.rs:235 on caller-provided input tensors: in = host f32
 * [batch,S,S,3]; regressors = host f32 [batch,N,16]; classificators = host f32 [batch,N,1]. */
FDL_API int fdl_detector_forward(fdl_detector*, const float* in, int batch, float* regressors, float* classificators);
/* decode_boxes + get_sigmoid_score + convert_to_detections + non_maximum_suppression +
 * detection_letterbox_removal (face_detection.rs:259-265) on caller-provided raw tensors.
 * padding = (left, top, right, bottom) per image (f64, types.rs:10).  Optional outputs, per
 * image with capacity cap_surv: survivor anchor indices in ascending order and, for each, the
 * index of the output detection whose cluster absorbed it. */
FDL_API int fdl_detector_postprocess(fdl_detector*, const float* regressors, const float* classificators, int batch,
                                     const double* padding4, fdl_detection* out, int cap_per_image, int* n_out,
                                     int32_t* survivor_anchor, int32_t* survivor_cluster, int cap_surv, int* n_surv);

/* ---------------------------------------------------------------- FaceLandmark */
/* FaceLandmark::new(model_path) face_landmark.rs:208-222: `model_file` is a FILE path
 * (NULL -> "./models/face_landmark.tflite"). */
FDL_API int fdl_landmark_create(const char* model_file, int device, fdl_landmark_model** out);
FDL_API void fdl_landmark_destroy(fdl_landmark_model*);
/* FaceLandmark::infer(&self, &Mat, Option<Rect>) face_landmark.rs:232-306.  *n_out = 468, or 0
 * when sigmoid(face flag) <= 0.5 (:292-296).  `out` must hold 468 entries.  `face_flag_logit`
 * (optional) receives the raw flag. */
FDL_API int fdl_landmark_infer(fdl_landmark_model*, const fdl_image* image, const fdl_rect* roi,
                               fdl_landmark* out, int* n_out, float* face_flag_logit);
/* invoke() face_landmark.rs:265: in [batch,192,192,3] -> landmarks [batch,1404], flag [batch,1]. */
FDL_API int fdl_landmark_forward(fdl_landmark_model*, const float* in, int batch, float* landmarks, float* flag);

/* ---------------------------------------------------------------- IrisLandmark */
/* IrisLandmark::new(model_path) iris_landmark.rs:142-156 (FILE path; NULL -> "./models/iris_landmark.tflite"). */
FDL_API int fdl_iris_create(const char* model_file, int device, fdl_iris_model** out);
FDL_API void fdl_iris_destroy(fdl_iris_model*);
/* IrisLandmark::infer(&self, &Mat, Option<Rect>, Option<bool>) iris_landmark.rs:158-248.
 * contour: 71 entries (the reference's IrisResults.contour), iris: 5 entries. */
FDL_API int fdl_iris_infer(fdl_iris_model*, const fdl_image* image, const fdl_rect* roi, int is_right_eye,
                           fdl_landmark* contour, fdl_landmark* iris);
/* invoke() iris_landmark.rs:203: in [batch,64,64,3] -> contours [batch,213], iris [batch,15]. */
FDL_API int fdl_iris_forward(fdl_iris_model*, const float* in, int batch, float* contours, float* iris);

/* ---------------------------------------------------------------- free functions */
/* face_detection_to_roi(Detection, (w,h), Option<SizeMode>) face_landmark.rs:180-198.
 * Evaluated on `device` by the same device function the pipeline uses. */
FDL_API int fdl_face_detection_to_roi(int device, const fdl_detection* det, int image_width, int image_height,
                                      int size_mode, fdl_rect* out);
/* iris_roi_from_face_landmarks(Vec<Landmark>, (w,h)) iris_landmark.rs:268-292; `n` must be 468
 * (the reference indexes 33/133/362/263 and panics on short input; here FDL_ERR_INVALID). */
FDL_API int fdl_iris_roi_from_face_landmarks(int device, const fdl_landmark* landmarks, int n, int image_width,
                                             int image_height, fdl_rect* left, fdl_rect* right);
/* image_to_tensor(...) transform.rs:188-309 (private in the reference; exported for parity
 * step 1).  out_tensor: host f32 [out_h,out_w,3]; out_u8 (optional): the uint8 image just before
 * normalisation; padding4: (left,top,right,bottom) f64. */
FDL_API int fdl_image_to_tensor(int device, const fdl_image* image, const fdl_rect* roi, int out_w, int out_h,
                                int keep_aspect_ratio, double range_min, double range_max, int flip_horizontal,
                                float* out_tensor, uint8_t* out_u8, double* padding4);
/* project_landmarks(...) transform.rs:351-432 (private in the reference; exported for parity
 * step 3). raw: host f32 [n*3]; out: n entries. roi may be NULL. */
FDL_API int fdl_project_landmarks(int device, const float* raw, int n, int tensor_w, int tensor_h, int image_w,
                                  int image_h, const double* padding4, const fdl_rect* roi, int flip_horizontal,
                                  fdl_landmark* out);

/* update_face_landmarks_with_iris_results(face_landmarks, iris_left, iris_right) iris_landmark.rs:380-398
 * (SURVEY.md 8f rank 1): contour point k of each eye replaces face landmark
 * LEFT_/RIGHT_EYE_TO_FACE_LANDMARK_INDEX[k] (:64-95).  n must be 468 ("unexpected number of items in
 * face_landmarks" -> FDL_ERR_INVALID); n_left / n_right <= 71.  refined: 468 entries. */
FDL_API int fdl_update_face_landmarks_with_iris_results(int device, const fdl_landmark* face_landmarks, int n,
                                                        const fdl_landmark* left_contour, int n_left,
                                                        const fdl_landmark* right_contour, int n_right, fdl_landmark* refined);
/* The two index maps themselves (host copy of the table the kernels use): out71[k] = face landmark index. */
FDL_API int fdl_eye_to_face_landmark_index(int is_right_eye, int32_t* out71);
/* get_iris_diameter(&iris_landmarks, image_size) iris_landmark.rs:401-418 (private in the reference): mean of the
 * Left-Right and Top-Bottom extents in pixels.  iris: 5 entries in IrisIndex order (:104-110). */
FDL_API int fdl_iris_diameter(int device, const fdl_landmark* iris, int n, int image_width, int image_height, double* diameter_px);
/* get_iris_depth(iris_landmarks, focal_length_mm, iris_size_px, image_size) iris_landmark.rs:421-433 (private, unused
 * in the reference): distance of the iris in mm from the 11.8 mm average human iris size. */
FDL_API int fdl_iris_depth(int device, const fdl_landmark* iris, int n, double focal_length_mm, double iris_size_px,
                           int image_width, int image_height, double* depth_mm);

/* ---------------------------------------------------------------- frame ingest */
/* utils.rs:8-21 `convert_image_to_mat(im_bytes)`: imgcodecs::imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB) -- here a baseline JPEG
 * decoder on the device (Huffman stage, jpeg_idct_islow, fancy chroma upsampling, fixed-point YCbCr -> RGB: libjpeg's defaults,
 * bit-exact with cv2.imdecode).  Accepted: 8-bit baseline (SOF0 / SOF1 Huffman) files with one interleaved scan, 1 or 3
 * components, luma at full resolution, chroma at 1x1, 2x1 or 2x2, with or without restart intervals, EXIF orientation 1 or
 * none.  Everything else (progressive, arithmetic, CMYK, rotated by EXIF, truncated scans) is FDL_ERR_INVALID with a message:
 * never a silent approximation, never a CPU fallback. */
/* Header only (host): size and component count.  No GPU needed. */
FDL_API int fdl_jpeg_info(const uint8_t* data, size_t len, int* width, int* height, int* components);
FDL_API int fdl_jpeg_decoder_create(int device, fdl_jpeg_decoder** out);
FDL_API void fdl_jpeg_decoder_destroy(fdl_jpeg_decoder*);
/* Decodes n files into tightly packed RGB images (rows of width*3 bytes) written back to back into `out` (host memory, or device
 * memory on the decoder's device when out_mem == FDL_MEM_DEVICE).  offsets[i] (optional) receives the byte offset of image i in
 * `out`, widths / heights (optional) its size.  FDL_ERR_CAPACITY when `cap` bytes do not hold the batch (offsets / sizes are
 * filled, so the caller can size the buffer from a first call with cap == 0). */
FDL_API int fdl_jpeg_decode(fdl_jpeg_decoder*, const uint8_t* const* data, const size_t* len, int n, uint8_t* out, size_t cap, int out_mem,
                            int64_t* offsets, int32_t* widths, int32_t* heights);
/* One-shot form of convert_image_to_mat: one file -> RGB in host memory (cap >= width*height*3; *width / *height always filled). */
FDL_API int fdl_decode_jpeg(int device, const uint8_t* data, size_t len, uint8_t* out_rgb, size_t cap, int* width, int* height);

/* ---------------------------------------------------------------- generic network handle */
/* A planned .tflite graph on one device (what replaces the TFLite interpreter). */
FDL_API int fdl_net_create(const char* tflite_file, int device, fdl_net** out);
FDL_API void fdl_net_destroy(fdl_net*);
FDL_API fdl_net* fdl_detector_net(fdl_detector*);
FDL_API fdl_net* fdl_landmark_net(fdl_landmark_model*);
FDL_API fdl_net* fdl_iris_net(fdl_iris_model*);
FDL_API int fdl_net_num_outputs(const fdl_net*);
/* Elements per batch item of the input (index -1) or of output `i`. */
FDL_API int64_t fdl_net_io_elems(const fdl_net*, int i);
/* Run on host tensors: in [batch, ...]; outs[i] host buffers of batch*fdl_net_io_elems(i). */
FDL_API int fdl_net_forward(fdl_net*, const float* in, int batch, float* const* outs, int n_outs);
/* Human-readable launch plan (one line per fused kernel step); returns bytes needed. Works
 * without a GPU when the handle was created with device = -1 (plan-only, cannot run). */
FDL_API int64_t fdl_net_describe(const fdl_net*, char* buf, int64_t cap);
/* Number of kernel launches of one forward pass. */
FDL_API int fdl_net_num_steps(const fdl_net*);
/* Select the arithmetic of the pointwise (1x1) contractions: 0 = fp32 FFMA (default for parity
 * checks), 1 = tensor-core split-TF32 (fp32-equivalent; the default), 2 = as 1 but
 * without the warp-specialised BlazeBlock kernel (every block on the serial tensor-core kernel;
 * bit-identical to 1, kept for A/B timing and as a race check). */
FDL_API int fdl_net_set_mode(fdl_net*, int mode);
/* Device-resident benchmark hook: run `iters` forward passes at `batch` on the net's own input
 * buffer (filled once from `in_or_null`, host f32, or left as is) and return the mean time per
 * pass in milliseconds measured with CUDA events on the net's stream. */
FDL_API int fdl_net_time_forward(fdl_net*, const float* in_or_null, int batch, int iters, float* ms_per_pass);
/* Per-launch timing inside whole forward passes: CUDA events recorded on the net's stream before
 * every planned step and after the last; ms_per_step[i] = mean duration of step i over `iters`
 * passes (cap >= fdl_net_num_steps).  bench.py uses it for the roofline of the dominant kernel. */
FDL_API int fdl_net_time_steps(fdl_net*, const float* in_or_null, int batch, int iters, float* ms_per_step, int cap);

/* ---------------------------------------------------------------- batched pipeline */
/* detect -> face ROI -> landmark -> eye ROIs -> iris(L,R) exactly as lib.rs:20-40, for a batch of
 * equally-sized frames, without leaving the device between stages. */
typedef struct fdl_pipeline_config {
  int32_t detector_model;   /* FDL_MODEL_* */
  int32_t device;
  int32_t max_batch;        /* frames per submit */
  int32_t max_faces;        /* faces per frame carried into landmark/iris (lib.rs uses faces[0]) */
  int32_t frame_width, frame_height;
  int32_t run_landmarks;    /* 0: detection only (BASELINE config 2/3) */
  int32_t run_iris;         /* 0: stop after landmarks (config 4) */
  const char* model_dir;    /* directory holding the .tflite files; NULL -> "./models" */
  int32_t zero_copy_host;   /* 1: contiguous PINNED host frames are read in place by the kernels (no whole-frame H2D copy) */
  int32_t refine_landmarks; /* 1: also fill fdl_face_result.refined_landmarks (update_face_landmarks_with_iris_results,
                               iris_landmark.rs:380-398) on the device; needs run_iris */
  double focal_length_mm;   /* > 0: also fill iris_depth_mm (get_iris_depth, iris_landmark.rs:421-433) */
} fdl_pipeline_config;

/* Per-face result record. */
typedef struct fdl_face_result {
  fdl_rect face_roi;                 /* face_detection_to_roi */
  float face_flag_logit;             /* raw conv2d_30 output */
  int32_t has_landmarks;             /* sigmoid(flag) > 0.5 */
  float landmarks[FDL_NUM_FACE_LANDMARKS * 3];      /* projected, normalised (values of Vec<Landmark>) */
  fdl_rect eye_roi[2];               /* [0] = left (33,133), [1] = right (362,263) */
  float eye_contour[2][FDL_NUM_EYE_CONTOUR * 3];
  float iris[2][FDL_NUM_IRIS * 3];
  /* iris refinement */
  float refined_landmarks[FDL_NUM_FACE_LANDMARKS * 3];   /* landmarks with both eye contours scattered in (cfg.refine_landmarks) */
  double iris_diameter_px[2];        /* get_iris_diameter per eye ([0] left, [1] right); 0 when the eye was not processed */
  double iris_depth_mm[2];           /* get_iris_depth per eye; 0 unless cfg.focal_length_mm > 0 */
} fdl_face_result;

typedef struct fdl_frame_result {
  int32_t n_detections;              /* entries of detections[]: min(n_total_detections, FDL_MAX_DETECTIONS) */
  int32_t n_faces;                   /* min(n_detections, max_faces) */
  int32_t n_total_detections;        /* detections weighted NMS produced for this frame; the reference returns an unbounded Vec
                                        (face_detection.rs:267) -- when this exceeds FDL_MAX_DETECTIONS the record is truncated
                                        and fdl_pipeline_collect reports FDL_ERR_CAPACITY (the results are still delivered) */
  fdl_detection detections[FDL_MAX_DETECTIONS];
} fdl_frame_result;

FDL_API int fdl_pipeline_create(const fdl_pipeline_config* cfg, fdl_pipeline** out);
FDL_API void fdl_pipeline_destroy(fdl_pipeline*);
/* Synchronous: frames[n] (host or device) -> frame_results[n], face_results[n*max_faces]. */
FDL_API int fdl_pipeline_run(fdl_pipeline*, const fdl_image* frames, int n, fdl_frame_result* frame_results,
                             fdl_face_result* face_results);
/* Asynchronous double-buffered form: submit enqueues H2D + all kernels + D2H on the pipeline's
 * streams and returns a ticket; collect waits for that ticket and copies the results out.  Up to
 * `fdl_pipeline_depth()` submits may be in flight.  collect (and run) return FDL_ERR_CAPACITY -- after
 * delivering the results and releasing the ticket -- when a frame produced more than FDL_MAX_DETECTIONS
 * detections (fdl_frame_result.n_total_detections says how many; fdl_detector_infer with a larger
 * buffer returns all of them), as fdl_detector_infer does for a short caller buffer. */
FDL_API int fdl_pipeline_depth(const fdl_pipeline*);
FDL_API int fdl_pipeline_submit(fdl_pipeline*, const fdl_image* frames, int n, int* ticket);
FDL_API int fdl_pipeline_collect(fdl_pipeline*, int ticket, fdl_frame_result* frame_results,
                                 fdl_face_result* face_results, int* n);
/* The lib.rs:20-40 sequence from where it really starts -- encoded bytes (`convert_image_to_mat`, utils.rs:8-21): n JPEG files
 * of the pipeline's frame size are copied to the device COMPRESSED (a 1080p frame is ~0.2 MB instead of 6.2 MB over PCIe),
 * decoded there (see fdl_jpeg_decode) into the lane's frame buffer and run through the same stages as fdl_pipeline_submit.
 * Pinned caller memory that holds the files close together in ascending order is read in place by the copy engine.  Errors of
 * the entropy-coded data (truncated scan, missing restart markers) surface from fdl_pipeline_collect as FDL_ERR_INVALID. */
FDL_API int fdl_pipeline_submit_jpeg(fdl_pipeline*, const uint8_t* const* data, const size_t* len, int n, int* ticket);
/* Device time of the last collected ticket's kernels (excludes copies), milliseconds. */
/* Host-only introspection of the zero-copy ingest plan (no GPU needed): which source rows of a
 * frame_width x frame_height frame the detector's letterbox (image_to_tensor with roi = None,
 * transform.rs:239-280, INTER_LINEAR 2x2 taps) reads for an input_size x input_size tensor, and how
 * the copy engine gathers them.  Returns 1 and fills row_pos[frame_height] (compact row index or
 * -1) and info4 = {compact rows per frame, source rows per period, periods per frame, strided
 * copies per batch} when the rows form a periodic pattern that continues across contiguous
 * frames; 0 when the pipeline reads the frames in place instead; negative on error. */
FDL_API int fdl_letterbox_row_plan(int frame_width, int frame_height, int input_size, int32_t* row_pos, int32_t* info4);
FDL_API float fdl_pipeline_last_device_ms(const fdl_pipeline*);
/* Stage breakdown of the last collected ticket, ms: [0] H2D (submit_jpeg: H2D of the compressed bytes + device decode), [1] detector preprocess, [2] detector
 * net, [3] SSD post-process, [4] face ROI + warp, [5] landmark net, [6] landmark post + eye warp,
 * [7] iris net, [8] iris post, [9] D2H. */
FDL_API int fdl_pipeline_stage_ms(const fdl_pipeline*, float* out10);

/* ---------------------------------------------------------------- drawing */
/* render.rs:361-479 `render_to_image(annotations, image, blend_mode)`: annotations painted over the frame in order (no blending --
 * the reference ignores blend_mode), with the pixel rules of the `imageproc` crate it calls.  An `Annotation` (render.rs:208-213:
 * data items + normalized_positions + thickness + colour) is flattened here into one fdl_primitive per data item:
 *   POINT        a, b = x, y                      -> filled square of side 2 * max(thickness / 2, 1) at (x - w, y - w)   (:421-430)
 *   LINE         a, b, c, d = x0, y0, x1, y1      -> one-pixel Bresenham segment between the truncated end points        (:431-442)
 *   RECT         a, b, c, d = left, top, right, bottom -> one-pixel outline of Rect::at(left, top).of_size(right - left, bottom - top) (:443-461)
 *   FILLED_RECT  the same rectangle, filled                                                                              (:462-474)
 * (ovals are drawn as rectangles by the reference.)  `normalized` coordinates are multiplied by the image size first.  The output is
 * RGBA8 [height][width][4] (DynamicImage::ImageRgba8), in host memory or -- out_mem == FDL_MEM_DEVICE -- on `device`.  The helpers
 * that build annotations from results (detections_to_render_data :262, landmarks_to_render_data :315, face_ / eye_landmarks_to_render_data,
 * the connection tables) are host-side bookkeeping and live in the mirrors (api.py, fdl.hpp). */
enum { FDL_PRIM_POINT = 0, FDL_PRIM_LINE = 1, FDL_PRIM_RECT = 2, FDL_PRIM_FILLED_RECT = 3 };
typedef struct fdl_primitive {
  int32_t kind;          /* FDL_PRIM_* */
  int32_t normalized;    /* Annotation.normalized_positions */
  double a, b, c, d;
  double thickness;      /* Annotation.thickness */
  uint8_t r, g, b_, alpha;   /* Annotation.color (`b_`: blue; alpha 255 when the reference's Option is None) */
  int32_t _pad;
} fdl_primitive;
FDL_API int fdl_render_to_image(int device, const fdl_image* image, const fdl_primitive* primitives, int n, uint8_t* out_rgba, size_t cap,
                                int out_mem);

/* ---------------------------------------------------------------- all GPUs of a box behind one handle */
/* The path shards by frame and has no exchange step (face_detection.rs:205 `infer(&self, ..)` is pure given the weights), so a box
 * of N GPUs is N independent pipelines.  fdl_pool owns one fdl_pipeline per listed device (a device may be listed more than once)
 * and one host worker thread per pipeline, pinned to the CPU cores local to that GPU: fdl_pool_submit* hands the batch to the
 * least-loaded device's worker and returns at once, so one application thread keeps every GPU fed -- header parsing, staging and
 * the ~300 stream operations of a batch run on the workers, in parallel.  Tickets are pool-wide.  `cfg->device` is ignored.
 * Frame memory given to fdl_pool_submit must stay valid until the ticket is collected; results come back in submit order per
 * ticket, exactly as from fdl_pipeline_collect (including FDL_ERR_CAPACITY / FDL_ERR_INVALID reports). */
FDL_API int fdl_pool_create(const fdl_pipeline_config* cfg, const int* devices, int n_devices, fdl_pool** out);
FDL_API void fdl_pool_destroy(fdl_pool*);
FDL_API int fdl_pool_devices(const fdl_pool*);             /* pipelines in the pool */
FDL_API int fdl_pool_depth(const fdl_pool*);               /* tickets that may be in flight: devices x fdl_pipeline_depth */
FDL_API int fdl_pool_submit(fdl_pool*, const fdl_image* frames, int n, int* ticket);
FDL_API int fdl_pool_submit_jpeg(fdl_pool*, const uint8_t* const* data, const size_t* len, int n, int* ticket);
/* device_index (optional) receives the index (into `devices`) of the pipeline that ran the ticket. */
FDL_API int fdl_pool_collect(fdl_pool*, int ticket, fdl_frame_result* frame_results, fdl_face_result* face_results, int* n,
                             int* device_index);

/* ---------------------------------------------------------------- a frame staged once for the per-frame API */
/* lib.rs:20-40 passes the same `&Mat` to FaceDetection::infer, FaceLandmark::infer and IrisLandmark::infer (x2).  With host images
 * every one of those calls uploads the frame again; an fdl_frame uploads it once (or decodes it on the device from JPEG bytes,
 * utils.rs:8-21) and fdl_frame_image() describes the device copy as an fdl_image (mem = FDL_MEM_DEVICE) that the infer calls
 * read in place.  The image stays valid until fdl_frame_destroy or the next upload into the same handle. */
FDL_API int fdl_frame_create(int device, fdl_frame** out);
FDL_API void fdl_frame_destroy(fdl_frame*);
FDL_API int fdl_frame_upload(fdl_frame*, const fdl_image* image);
FDL_API int fdl_frame_upload_jpeg(fdl_frame*, const uint8_t* data, size_t len);
FDL_API int fdl_frame_image(const fdl_frame*, fdl_image* out);

#ifdef __cplusplus
}
#endif
#endif /* FDL_H_ */
