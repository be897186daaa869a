// fdl.hpp -- header-only C++17 mirror of the reference's public API over the C ABI (fdl.h).
// Same names and argument meaning as face_detection_lite::{FaceDetection, FaceLandmark, IrisLandmark,
// face_detection_to_roi, iris_roi_from_face_landmarks}; where the reference returns Err this throws fdl::Error.
#pragma once
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fdl.h"

namespace fdl {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != FDL_OK) throw Error(rc, fdl_last_error()); }

using Rect = fdl_rect;            // types.rs:24-37
using Landmark = fdl_landmark;    // types.rs:176-187
struct Detection {                // types.rs:189-246 (data = [8,2] row-major)
  float data[16]; float score; int anchor;
};
enum class FaceDetectionModel { FrontCamera = 0, BackCamera = 1, Short = 2, Full = 3, FullSparse = 4 };   // face_detection.rs:117-123
enum class SizeMode { Default = 0, SquareLong = 1, SquareShort = 2 };                                      // transform.rs:15-24

// The `&Mat` argument: 8UC3 RGB, HWC.
struct Image {
  const uint8_t* data; int width, height; long long row_stride = 0; bool on_device = false;
  fdl_image c() const { return fdl_image{data, width, height, row_stride, on_device ? FDL_MEM_DEVICE : FDL_MEM_HOST, 0}; }
};

class FaceDetection {             // face_detection.rs:146-267
 public:
  FaceDetection(FaceDetectionModel model, std::optional<std::string> model_path = std::nullopt, int device = 0) {
    check(fdl_detector_create((int)model, model_path ? model_path->c_str() : nullptr, device, &h_));
  }
  ~FaceDetection() { fdl_detector_destroy(h_); }
  FaceDetection(const FaceDetection&) = delete;
  std::vector<Detection> infer(const Image& image, std::optional<Rect> roi = std::nullopt, int max_detections = 128) {
    std::vector<fdl_detection> out(max_detections);
    int n = 0;
    fdl_image img = image.c();
    check(fdl_detector_infer(h_, &img, roi ? &*roi : nullptr, out.data(), max_detections, &n));
    std::vector<Detection> r(n);
    for (int i = 0; i < n; ++i) { for (int k = 0; k < 16; ++k) r[i].data[k] = out[i].data[k]; r[i].score = out[i].score; r[i].anchor = out[i].anchor; }
    return r;
  }
 private:
  fdl_detector* h_ = nullptr;
};

inline Rect face_detection_to_roi(const Detection& d, std::pair<int, int> image_size, std::optional<SizeMode> mode = std::nullopt, int device = 0) {
  fdl_detection c{}; for (int k = 0; k < 16; ++k) c.data[k] = d.data[k]; c.score = d.score; c.anchor = d.anchor;
  Rect out{};
  check(fdl_face_detection_to_roi(device, &c, image_size.first, image_size.second, mode ? (int)*mode : FDL_SIZE_MODE_NONE, &out));   // face_landmark.rs:180
  return out;
}

class FaceLandmark {              // face_landmark.rs:200-306
 public:
  explicit FaceLandmark(std::optional<std::string> model_path = std::nullopt, int device = 0) {
    check(fdl_landmark_create(model_path ? model_path->c_str() : nullptr, device, &h_));
  }
  ~FaceLandmark() { fdl_landmark_destroy(h_); }
  FaceLandmark(const FaceLandmark&) = delete;
  std::vector<Landmark> infer(const Image& image, std::optional<Rect> roi = std::nullopt) {
    std::vector<Landmark> out(FDL_NUM_FACE_LANDMARKS);
    int n = 0; float flag = 0.f;
    fdl_image img = image.c();
    check(fdl_landmark_infer(h_, &img, roi ? &*roi : nullptr, out.data(), &n, &flag));
    out.resize(n);
    return out;
  }
 private:
  fdl_landmark_model* h_ = nullptr;
};

inline std::pair<Rect, Rect> iris_roi_from_face_landmarks(const std::vector<Landmark>& lm, std::pair<int, int> image_size, int device = 0) {
  Rect l{}, r{};
  check(fdl_iris_roi_from_face_landmarks(device, lm.data(), (int)lm.size(), image_size.first, image_size.second, &l, &r));   // iris_landmark.rs:268
  return {l, r};
}

struct IrisResults {              // iris_landmark.rs:115-129
  std::vector<Landmark> contour, iris;
  std::vector<Landmark> eyeball_contour() const { return {contour.begin(), contour.begin() + 15}; }
};

// update_face_landmarks_with_iris_results (iris_landmark.rs:380-398)
inline std::vector<Landmark> update_face_landmarks_with_iris_results(const std::vector<Landmark>& face_landmarks, const IrisResults& iris_data_left,
                                                                     const IrisResults& iris_data_right, int device = 0) {
  std::vector<Landmark> out(FDL_NUM_FACE_LANDMARKS);
  check(fdl_update_face_landmarks_with_iris_results(device, face_landmarks.data(), (int)face_landmarks.size(), iris_data_left.contour.data(),
                                                    (int)iris_data_left.contour.size(), iris_data_right.contour.data(),
                                                    (int)iris_data_right.contour.size(), out.data()));
  return out;
}
// get_iris_diameter / get_iris_depth (iris_landmark.rs:401-433; private in the reference)
inline double get_iris_diameter(const std::vector<Landmark>& iris_landmarks, std::pair<int, int> image_size, int device = 0) {
  double d = 0.0;
  check(fdl_iris_diameter(device, iris_landmarks.data(), (int)iris_landmarks.size(), image_size.first, image_size.second, &d));
  return d;
}
inline double get_iris_depth(const std::vector<Landmark>& iris_landmarks, double focal_length_mm, double iris_size_px, std::pair<int, int> image_size,
                             int device = 0) {
  double d = 0.0;
  check(fdl_iris_depth(device, iris_landmarks.data(), (int)iris_landmarks.size(), focal_length_mm, iris_size_px, image_size.first, image_size.second, &d));
  return d;
}

class IrisLandmark {              // iris_landmark.rs:131-248
 public:
  explicit IrisLandmark(std::optional<std::string> model_path = std::nullopt, int device = 0) {
    check(fdl_iris_create(model_path ? model_path->c_str() : nullptr, device, &h_));
  }
  ~IrisLandmark() { fdl_iris_destroy(h_); }
  IrisLandmark(const IrisLandmark&) = delete;
  IrisResults infer(const Image& image, std::optional<Rect> roi = std::nullopt, std::optional<bool> is_right_eye = std::nullopt) {
    IrisResults r;
    r.contour.resize(FDL_NUM_EYE_CONTOUR); r.iris.resize(FDL_NUM_IRIS);
    fdl_image img = image.c();
    check(fdl_iris_infer(h_, &img, roi ? &*roi : nullptr, is_right_eye.value_or(false) ? 1 : 0, r.contour.data(), r.iris.data()));
    return r;
  }
 private:
  fdl_iris_model* h_ = nullptr;
};

// render.rs: Color (:7-27), Annotation (:208-213, flattened: every item carries its kind) and render_to_image (:361-479) -> RGBA8.
struct Color { int r = 0, g = 0, b = 0; std::optional<int> a; };
struct AnnotationItem { int kind; double a, b, c, d; std::optional<Color> fill; };   // kind: FDL_PRIM_*; fill: FilledRectOrOval's own colour
struct Annotation { std::vector<AnnotationItem> data; bool normalized_positions = true; double thickness = 1.0; Color color; };
inline Annotation landmark_points(const std::vector<Landmark>& lm, Color color, double thickness = 2.0) {   // landmarks_to_render_data, the points
  Annotation a{{}, true, thickness, color};
  for (const auto& l : lm) a.data.push_back({FDL_PRIM_POINT, l.x, l.y, 0.0, 0.0, std::nullopt});
  return a;
}
inline std::vector<uint8_t> render_to_image(const std::vector<Annotation>& annotations, const Image& image, int device = 0) {
  std::vector<fdl_primitive> prims;
  for (const auto& an : annotations)
    for (const auto& it : an.data) {
      const Color& c = it.fill ? *it.fill : an.color;
      prims.push_back(fdl_primitive{it.kind, an.normalized_positions ? 1 : 0, it.a, it.b, it.c, it.d, an.thickness, (uint8_t)c.r, (uint8_t)c.g, (uint8_t)c.b,
                                    (uint8_t)c.a.value_or(255), 0});
    }
  std::vector<uint8_t> out((size_t)image.width * image.height * 4);
  fdl_image img = image.c();
  check(fdl_render_to_image(device, &img, prims.data(), (int)prims.size(), out.data(), out.size(), FDL_MEM_HOST));
  return out;
}

}  // namespace fdl
