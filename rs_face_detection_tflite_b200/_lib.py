"""ctypes view of include/fdl.h.  Loads the in-tree libfdl_b200.so and fails loudly when it is missing:
there is no Python / CPU fallback for any call in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FDL_LIB: an alternative build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("FDL_LIB") or os.path.join(HERE, "libfdl_b200.so")

FDL_OK = 0
FDL_ERR_INVALID, FDL_ERR_IO, FDL_ERR_MODEL, FDL_ERR_CUDA, FDL_ERR_CAPACITY, FDL_ERR_INTERNAL = -1, -2, -3, -4, -5, -6
NUM_FACE_LANDMARKS, NUM_EYE_CONTOUR, NUM_IRIS, MAX_DETECTIONS = 468, 71, 5, 32
MEM_HOST, MEM_DEVICE = 0, 1


class FdlError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("fdl error %d: %s" % (code, message))
        self.code = code
        self.message = message


class CRect(C.Structure):
    _fields_ = [("x_center", C.c_double), ("y_center", C.c_double), ("width", C.c_double), ("height", C.c_double),
                ("rotation", C.c_double), ("normalized", C.c_int32), ("_pad", C.c_int32)]


class CDetection(C.Structure):
    _fields_ = [("data", C.c_float * 16), ("score", C.c_float), ("anchor", C.c_int32)]


class CLandmark(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class CImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_stride", C.c_int64),
                ("mem", C.c_int32), ("_pad", C.c_int32)]


class CPipelineConfig(C.Structure):
    _fields_ = [("detector_model", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32), ("max_faces", C.c_int32),
                ("frame_width", C.c_int32), ("frame_height", C.c_int32), ("run_landmarks", C.c_int32), ("run_iris", C.c_int32),
                ("model_dir", C.c_char_p), ("zero_copy_host", C.c_int32), ("refine_landmarks", C.c_int32),
                ("focal_length_mm", C.c_double)]


class CFaceResult(C.Structure):
    _fields_ = [("face_roi", CRect), ("face_flag_logit", C.c_float), ("has_landmarks", C.c_int32),
                ("landmarks", C.c_float * (NUM_FACE_LANDMARKS * 3)), ("eye_roi", CRect * 2),
                ("eye_contour", (C.c_float * (NUM_EYE_CONTOUR * 3)) * 2), ("iris", (C.c_float * (NUM_IRIS * 3)) * 2),
                ("refined_landmarks", C.c_float * (NUM_FACE_LANDMARKS * 3)), ("iris_diameter_px", C.c_double * 2),
                ("iris_depth_mm", C.c_double * 2)]


class CPrimitive(C.Structure):
    _fields_ = [("kind", C.c_int32), ("normalized", C.c_int32), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("d", C.c_double),
                ("thickness", C.c_double), ("r", C.c_uint8), ("g", C.c_uint8), ("b_", C.c_uint8), ("alpha", C.c_uint8), ("_pad", C.c_int32)]


class CFrameResult(C.Structure):
    _fields_ = [("n_detections", C.c_int32), ("n_faces", C.c_int32), ("n_total_detections", C.c_int32),
                ("detections", CDetection * MAX_DETECTIONS)]


# every symbol include/fdl.h declares: name -> (restype, argtypes)
_P = C.POINTER
_vp = C.c_void_p
SYMBOLS = {
    "fdl_last_error": (C.c_char_p, []),
    "fdl_version": (C.c_char_p, []),
    "fdl_device_count": (C.c_int, []),
    "fdl_launch_count": (C.c_uint64, []),
    "fdl_detector_create": (C.c_int, [C.c_int, C.c_char_p, C.c_int, _P(_vp)]),
    "fdl_detector_destroy": (None, [_vp]),
    "fdl_detector_input_size": (C.c_int, [_vp]),
    "fdl_detector_num_anchors": (C.c_int, [_vp]),
    "fdl_detector_anchors": (C.c_int, [_vp, _P(C.c_float), C.c_int]),
    "fdl_detector_infer": (C.c_int, [_vp, _P(CImage), _P(CRect), _P(CDetection), C.c_int, _P(C.c_int)]),
    "fdl_detector_infer_batch": (C.c_int, [_vp, _P(CImage), C.c_int, _P(CDetection), C.c_int, _P(C.c_int)]),
    "fdl_detector_forward": (C.c_int, [_vp, _P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float)]),
    "fdl_detector_postprocess": (C.c_int, [_vp, _P(C.c_float), _P(C.c_float), C.c_int, _P(C.c_double), _P(CDetection), C.c_int,
                                           _P(C.c_int), _P(C.c_int32), _P(C.c_int32), C.c_int, _P(C.c_int)]),
    "fdl_landmark_create": (C.c_int, [C.c_char_p, C.c_int, _P(_vp)]),
    "fdl_landmark_destroy": (None, [_vp]),
    "fdl_landmark_infer": (C.c_int, [_vp, _P(CImage), _P(CRect), _P(CLandmark), _P(C.c_int), _P(C.c_float)]),
    "fdl_landmark_forward": (C.c_int, [_vp, _P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float)]),
    "fdl_iris_create": (C.c_int, [C.c_char_p, C.c_int, _P(_vp)]),
    "fdl_iris_destroy": (None, [_vp]),
    "fdl_iris_infer": (C.c_int, [_vp, _P(CImage), _P(CRect), C.c_int, _P(CLandmark), _P(CLandmark)]),
    "fdl_iris_forward": (C.c_int, [_vp, _P(C.c_float), C.c_int, _P(C.c_float), _P(C.c_float)]),
    "fdl_face_detection_to_roi": (C.c_int, [C.c_int, _P(CDetection), C.c_int, C.c_int, C.c_int, _P(CRect)]),
    "fdl_iris_roi_from_face_landmarks": (C.c_int, [C.c_int, _P(CLandmark), C.c_int, C.c_int, C.c_int, _P(CRect), _P(CRect)]),
    "fdl_image_to_tensor": (C.c_int, [C.c_int, _P(CImage), _P(CRect), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                      _P(C.c_float), _P(C.c_uint8), _P(C.c_double)]),
    "fdl_project_landmarks": (C.c_int, [C.c_int, _P(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_double), _P(CRect),
                                        C.c_int, _P(CLandmark)]),
    "fdl_update_face_landmarks_with_iris_results": (C.c_int, [C.c_int, _P(CLandmark), C.c_int, _P(CLandmark), C.c_int, _P(CLandmark), C.c_int,
                                                              _P(CLandmark)]),
    "fdl_eye_to_face_landmark_index": (C.c_int, [C.c_int, _P(C.c_int32)]),
    "fdl_iris_diameter": (C.c_int, [C.c_int, _P(CLandmark), C.c_int, C.c_int, C.c_int, _P(C.c_double)]),
    "fdl_iris_depth": (C.c_int, [C.c_int, _P(CLandmark), C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _P(C.c_double)]),
    "fdl_net_create": (C.c_int, [C.c_char_p, C.c_int, _P(_vp)]),
    "fdl_net_destroy": (None, [_vp]),
    "fdl_detector_net": (_vp, [_vp]),
    "fdl_landmark_net": (_vp, [_vp]),
    "fdl_iris_net": (_vp, [_vp]),
    "fdl_net_num_outputs": (C.c_int, [_vp]),
    "fdl_net_io_elems": (C.c_int64, [_vp, C.c_int]),
    "fdl_net_forward": (C.c_int, [_vp, _P(C.c_float), C.c_int, _P(_P(C.c_float)), C.c_int]),
    "fdl_net_describe": (C.c_int64, [_vp, C.c_char_p, C.c_int64]),
    "fdl_net_num_steps": (C.c_int, [_vp]),
    "fdl_net_set_mode": (C.c_int, [_vp, C.c_int]),
    "fdl_net_time_forward": (C.c_int, [_vp, _P(C.c_float), C.c_int, C.c_int, _P(C.c_float)]),
    "fdl_net_time_steps": (C.c_int, [_vp, _P(C.c_float), C.c_int, C.c_int, _P(C.c_float), C.c_int]),
    "fdl_letterbox_row_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_int32)]),
    "fdl_pipeline_create": (C.c_int, [_P(CPipelineConfig), _P(_vp)]),
    "fdl_pipeline_destroy": (None, [_vp]),
    "fdl_pipeline_run": (C.c_int, [_vp, _P(CImage), C.c_int, _P(CFrameResult), _P(CFaceResult)]),
    "fdl_pipeline_depth": (C.c_int, [_vp]),
    "fdl_pipeline_submit": (C.c_int, [_vp, _P(CImage), C.c_int, _P(C.c_int)]),
    "fdl_pipeline_collect": (C.c_int, [_vp, C.c_int, _P(CFrameResult), _P(CFaceResult), _P(C.c_int)]),
    "fdl_pipeline_submit_jpeg": (C.c_int, [_vp, _P(_vp), _P(C.c_size_t), C.c_int, _P(C.c_int)]),
    "fdl_jpeg_info": (C.c_int, [_vp, C.c_size_t, _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "fdl_jpeg_decoder_create": (C.c_int, [C.c_int, _P(_vp)]),
    "fdl_jpeg_decoder_destroy": (None, [_vp]),
    "fdl_jpeg_decode": (C.c_int, [_vp, _P(_vp), _P(C.c_size_t), C.c_int, _vp, C.c_size_t, C.c_int, _P(C.c_int64), _P(C.c_int32), _P(C.c_int32)]),
    "fdl_decode_jpeg": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, C.c_size_t, _P(C.c_int), _P(C.c_int)]),
    "fdl_render_to_image": (C.c_int, [C.c_int, _P(CImage), _P(CPrimitive), C.c_int, _vp, C.c_size_t, C.c_int]),
    "fdl_pool_create": (C.c_int, [_P(CPipelineConfig), _P(C.c_int), C.c_int, _P(_vp)]),
    "fdl_pool_destroy": (None, [_vp]),
    "fdl_pool_devices": (C.c_int, [_vp]),
    "fdl_pool_depth": (C.c_int, [_vp]),
    "fdl_pool_submit": (C.c_int, [_vp, _P(CImage), C.c_int, _P(C.c_int)]),
    "fdl_pool_submit_jpeg": (C.c_int, [_vp, _P(_vp), _P(C.c_size_t), C.c_int, _P(C.c_int)]),
    "fdl_pool_collect": (C.c_int, [_vp, C.c_int, _P(CFrameResult), _P(CFaceResult), _P(C.c_int), _P(C.c_int)]),
    "fdl_frame_create": (C.c_int, [C.c_int, _P(_vp)]),
    "fdl_frame_destroy": (None, [_vp]),
    "fdl_frame_upload": (C.c_int, [_vp, _P(CImage)]),
    "fdl_frame_upload_jpeg": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fdl_frame_image": (C.c_int, [_vp, _P(CImage)]),
    "fdl_pipeline_last_device_ms": (C.c_float, [_vp]),
    "fdl_pipeline_stage_ms": (C.c_int, [_vp, _P(C.c_float)]),
}

_lib = None


def lib():
    """The loaded library; raises if libfdl_b200.so has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FdlError(FDL_ERR_INTERNAL, "libfdl_b200.so is missing at %s: build it with "
                           "`python -m rs_face_detection_tflite_b200.build` (there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError here means the .so does not match include/fdl.h
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != FDL_OK:
        raise FdlError(rc, (lib().fdl_last_error() or b"").decode("utf-8", "replace"))
