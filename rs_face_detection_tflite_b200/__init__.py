"""B200-native detect -> landmark -> iris path of rs-face-detection-tflite.

Python mirror (ctypes) of the reference's public API over the C ABI in include/fdl.h; all compute is in
libfdl_b200.so (hand-written sm_100a CUDA).  See DESIGN.md / INTEGRATION.md.
"""
from .api import (BBox, Detection, FaceDetection, FaceDetectionModel, FaceIndex, FaceLandmark, FaceResult, FrameResult, IrisLandmark,
                  IrisResults, Landmark, Net, Pipeline, Rect, SizeMode, device_count, face_detection_to_roi, image_to_tensor,
                  iris_roi_from_face_landmarks, launch_count, letterbox_row_plan, project_landmarks, eye_to_face_landmark_index,
                  get_iris_depth, get_iris_diameter, update_face_landmarks_with_iris_results, JpegDecoder, convert_image_to_mat, jpeg_info, Frame, Pool, Annotation, Color, Colors, detections_to_render_data,
                  landmarks_to_render_data, face_landmarks_to_render_data, eye_landmarks_to_render_data, iris_landmarks_to_render_data, render_to_image,
                  FACE_LANDMARK_CONNECTIONS, EYE_LANDMARK_CONNECTIONS)
from ._lib import FdlError

__all__ = ["BBox", "Detection", "FaceDetection", "FaceDetectionModel", "FaceIndex", "FaceLandmark", "FaceResult", "FrameResult",
           "IrisLandmark", "IrisResults", "Landmark", "Net", "Pipeline", "Rect", "SizeMode", "FdlError", "device_count",
           "face_detection_to_roi", "image_to_tensor", "iris_roi_from_face_landmarks", "launch_count", "letterbox_row_plan", "project_landmarks",
           "eye_to_face_landmark_index", "get_iris_depth", "get_iris_diameter", "update_face_landmarks_with_iris_results",
           "JpegDecoder", "convert_image_to_mat", "jpeg_info", "Frame", "Pool", "Annotation", "Color", "Colors", "detections_to_render_data",
           "landmarks_to_render_data", "face_landmarks_to_render_data", "eye_landmarks_to_render_data", "iris_landmarks_to_render_data", "render_to_image",
           "FACE_LANDMARK_CONNECTIONS", "EYE_LANDMARK_CONNECTIONS"]
