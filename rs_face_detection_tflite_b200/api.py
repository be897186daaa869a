"""Host-side mirror of the reference's public API over the C ABI (include/fdl.h).

Names, argument meaning and error behaviour follow
``face_detection_lite::{face_detection, face_landmark, iris_landmark, types}`` of the reference
(paths relative to /root/reference/src/face_detection_lite/):

* ``FaceDetectionModel``                       face_detection.rs:117-123
* ``FaceDetection(model_type, model_path)``    face_detection.rs:153  (``model_path`` is a DIRECTORY)
* ``FaceDetection.infer(image, roi)``          face_detection.rs:205
* ``FaceLandmark(model_path).infer(image, roi)``          face_landmark.rs:208, :232  (FILE path)
* ``IrisLandmark(model_path).infer(image, roi, is_right_eye)``  iris_landmark.rs:142, :158
* ``face_detection_to_roi(detection, image_size, size_mode)``   face_landmark.rs:180
* ``iris_roi_from_face_landmarks(landmarks, image_size)``       iris_landmark.rs:268
* ``Rect``, ``BBox``, ``Detection``, ``Landmark``, ``IrisResults``  types.rs, iris_landmark.rs:115-129

Images are ``numpy`` uint8 arrays ``[H, W, 3]`` in RGB order (the ``Mat`` the reference gets from
``convert_image_to_mat``), or CUDA ``torch`` uint8 tensors of the same shape (no host copy).  Where the
reference returns ``Err`` this module raises ``FdlError``.  Everything is computed by libfdl_b200.so on
the GPU; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (CDetection, CFaceResult, CFrameResult, CImage, CLandmark, CPipelineConfig, CRect, FdlError, check, lib)

NUM_LANDMARKS = 468          # face_landmark.rs:27
ROI_SCALE = (1.5, 1.5)       # face_landmark.rs:30
IRIS_ROI_SCALE = (2.3, 2.3)  # iris_landmark.rs:27


class FaceDetectionModel(enum.IntEnum):  # face_detection.rs:117-123
    FrontCamera = 0
    BackCamera = 1
    Short = 2
    Full = 3
    FullSparse = 4


class SizeMode(enum.IntEnum):  # transform.rs:15-24
    Default = 0
    SquareLong = 1
    SquareShort = 2


class FaceIndex(enum.IntEnum):  # face_detection.rs:89-98
    LeftEye = 0
    RightEye = 1
    NoseTip = 2
    Mouth = 3
    LeftEyeTragion = 4
    RightEyeTragion = 5


@dataclass
class Rect:  # types.rs:24-37
    x_center: float
    y_center: float
    width: float
    height: float
    rotation: float
    normalized: bool

    def _c(self) -> CRect:
        return CRect(self.x_center, self.y_center, self.width, self.height, self.rotation, 1 if self.normalized else 0, 0)

    @staticmethod
    def _from(c: CRect) -> "Rect":
        return Rect(c.x_center, c.y_center, c.width, c.height, c.rotation, bool(c.normalized))


@dataclass
class BBox:  # types.rs:99-174
    xmin: float
    ymin: float
    xmax: float
    ymax: float

    @property
    def width(self):
        return self.xmax - self.xmin

    @property
    def height(self):
        return self.ymax - self.ymin


@dataclass
class Landmark:  # types.rs:176-187
    x: float
    y: float
    z: float


@dataclass
class Detection:  # types.rs:189-246
    data: np.ndarray      # [8,2] float32: row0 (xmin,ymin), row1 (xmax,ymax), rows 2..7 keypoints
    score: float
    anchor: int = -1      # extension: SSD anchor index of the NMS cluster's top detection

    def bbox(self) -> BBox:  # :215-221
        d = self.data
        return BBox(float(d[0, 0]), float(d[0, 1]), float(d[1, 0]), float(d[1, 1]))

    def keypoint(self, k: int):  # :209-212
        return float(self.data[k + 2, 0]), float(self.data[k + 2, 1])

    def _c(self) -> CDetection:
        c = CDetection()
        flat = np.ascontiguousarray(self.data, np.float32).reshape(16)
        for i in range(16):
            c.data[i] = float(flat[i])
        c.score = float(self.score)
        c.anchor = int(self.anchor)
        return c

    @staticmethod
    def _from(c: CDetection) -> "Detection":
        return Detection(np.array(c.data[:], np.float32).reshape(8, 2), float(np.float32(c.score)), int(c.anchor))


@dataclass
class IrisResults:  # iris_landmark.rs:115-129
    contour: list   # 71 Landmark
    iris: list      # 5 Landmark

    def eyeball_contour(self):  # iris_landmark.rs:126-128: first 15 contour points
        return self.contour[:15]


# ------------------------------------------------------------------------------------------------
class Frame:
    """A frame staged once on the device for the per-frame API (fdl_frame): lib.rs:20-40 hands the same ``&Mat`` to four ``infer``
    calls; with a ``Frame`` the pixels cross PCIe once -- or never uncompressed: ``Frame(jpeg=bytes)`` is ``convert_image_to_mat``
    (utils.rs:8-21) decoded on the device.  Pass it wherever an image is accepted."""

    def __init__(self, image=None, jpeg=None, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fdl_frame_create(device, C.byref(self._h)))
        self.device = device
        if image is not None:
            self.upload(image)
        elif jpeg is not None:
            self.upload_jpeg(jpeg)

    def upload(self, image):
        img, _keep = _image(image)
        check(lib().fdl_frame_upload(self._h, C.byref(img)))
        return self

    def upload_jpeg(self, data):
        a, ln, _k = _buffer_address(data)
        check(lib().fdl_frame_upload_jpeg(self._h, a, ln))
        return self

    def _cimage(self) -> CImage:
        img = CImage()
        check(lib().fdl_frame_image(self._h, C.byref(img)))
        return img

    @property
    def size(self):
        """(width, height), as ``Mat::size()``."""
        img = self._cimage()
        return img.width, img.height

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_frame_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _image(image):
    """-> (CImage, keepalive).  numpy uint8 [H,W,3] (host), CUDA torch uint8 tensor [H,W,3], or a Frame."""
    if isinstance(image, Frame):
        return image._cimage(), image
    if isinstance(image, np.ndarray):
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise FdlError(_lib.FDL_ERR_INVALID, "image must be uint8 [H,W,3] RGB")
        if image.strides[2] != 1 or image.strides[1] != 3:
            image = np.ascontiguousarray(image)
        return CImage(image.ctypes.data, image.shape[1], image.shape[0], image.strides[0], _lib.MEM_HOST, 0), image
    # torch tensor (host pinned or CUDA)
    t = image
    if t.dim() != 3 or t.shape[2] != 3 or str(t.dtype) != "torch.uint8":
        raise FdlError(_lib.FDL_ERR_INVALID, "image must be uint8 [H,W,3] RGB")
    if not t.is_contiguous():
        t = t.contiguous()
    mem = _lib.MEM_DEVICE if t.is_cuda else _lib.MEM_HOST
    return CImage(t.data_ptr(), t.shape[1], t.shape[0], t.shape[1] * 3, mem, 0), t


def _landmarks(buf, n):
    return [Landmark(buf[i].x, buf[i].y, buf[i].z) for i in range(n)]


class _Net:
    """Introspection / stage-level access to a planned graph (parity protocol step 2)."""

    def __init__(self, handle):
        self._h = handle

    def describe(self) -> str:
        n = lib().fdl_net_describe(self._h, None, 0)
        buf = C.create_string_buffer(int(n))
        lib().fdl_net_describe(self._h, buf, n)
        return buf.value.decode()

    @property
    def num_steps(self):
        return lib().fdl_net_num_steps(self._h)

    def set_mode(self, mode: int):
        check(lib().fdl_net_set_mode(self._h, mode))

    def forward(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        b = x.shape[0]
        assert x[0].size == lib().fdl_net_io_elems(self._h, -1), "input shape does not match the graph"
        n_out = lib().fdl_net_num_outputs(self._h)
        outs = [np.empty((b, lib().fdl_net_io_elems(self._h, i)), np.float32) for i in range(n_out)]
        ptrs = (C.POINTER(C.c_float) * n_out)(*[o.ctypes.data_as(C.POINTER(C.c_float)) for o in outs])
        check(lib().fdl_net_forward(self._h, x.ctypes.data_as(C.POINTER(C.c_float)), b, ptrs, n_out))
        return outs

    def time_forward(self, batch: int, iters: int, x: np.ndarray | None = None) -> float:
        ms = C.c_float()
        p = None
        if x is not None:
            x = np.ascontiguousarray(x, np.float32)
            p = x.ctypes.data_as(C.POINTER(C.c_float))
        check(lib().fdl_net_time_forward(self._h, p, batch, iters, C.byref(ms)))
        return ms.value

    def time_steps(self, batch: int, iters: int, x: np.ndarray | None = None) -> np.ndarray:
        """Mean duration (ms) of every planned launch inside whole forward passes (CUDA events on the net's stream)."""
        n = self.num_steps
        out = (C.c_float * n)()
        p = None
        if x is not None:
            x = np.ascontiguousarray(x, np.float32)
            p = x.ctypes.data_as(C.POINTER(C.c_float))
        check(lib().fdl_net_time_steps(self._h, p, batch, iters, out, n))
        return np.array(out[:], np.float32)


class Net(_Net):
    """A stand-alone planned .tflite graph (``device=-1``: plan only, no GPU needed)."""

    def __init__(self, tflite_file: str, device: int = 0):
        h = C.c_void_p()
        check(lib().fdl_net_create(os.fsencode(tflite_file), device, C.byref(h)))
        super().__init__(h)

    def close(self):
        if self._h:
            lib().fdl_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
class FaceDetection:
    """BlazeFace detector.  ``FaceDetection::new`` / ``infer`` (face_detection.rs:153, :205)."""

    def __init__(self, model_type: FaceDetectionModel = FaceDetectionModel.FrontCamera, model_path: str | None = None, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fdl_detector_create(int(model_type), os.fsencode(model_path) if model_path else None, device, C.byref(self._h)))
        self.model_type = FaceDetectionModel(int(model_type))
        self.device = device
        self.input_size = lib().fdl_detector_input_size(self._h)
        self.num_anchors = lib().fdl_detector_num_anchors(self._h)
        self.net = _Net(lib().fdl_detector_net(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def anchors(self) -> np.ndarray:
        """ssd_generate_anchors (face_detection.rs:366-413): [N,2] float32."""
        a = np.empty((self.num_anchors, 2), np.float32)
        check(lib().fdl_detector_anchors(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), self.num_anchors))
        return a

    def infer(self, image, roi: Rect | None = None, max_detections: int = 64):
        img, _keep = _image(image)
        out = (CDetection * max_detections)()
        n = C.c_int()
        croi = roi._c() if roi is not None else None
        check(lib().fdl_detector_infer(self._h, C.byref(img), C.byref(croi) if croi is not None else None, out, max_detections, C.byref(n)))
        return [Detection._from(out[i]) for i in range(n.value)]

    def infer_batch(self, images, max_detections: int = 64):
        imgs = [_image(i) for i in images]
        arr = (CImage * len(imgs))(*[i[0] for i in imgs])
        out = (CDetection * (max_detections * len(imgs)))()
        n = (C.c_int * len(imgs))()
        check(lib().fdl_detector_infer_batch(self._h, arr, len(imgs), out, max_detections, n))
        return [[Detection._from(out[b * max_detections + i]) for i in range(n[b])] for b in range(len(imgs))]

    # -- stage-level hooks (parity protocol) --
    def forward(self, tensor: np.ndarray):
        """interpreter.invoke() on caller tensors: [B,S,S,3] -> (regressors [B,N,16], classificators [B,N,1])."""
        x = np.ascontiguousarray(tensor, np.float32)
        b = x.shape[0]
        reg = np.empty((b, self.num_anchors, 16), np.float32)
        cls = np.empty((b, self.num_anchors, 1), np.float32)
        fp = C.POINTER(C.c_float)
        check(lib().fdl_detector_forward(self._h, x.ctypes.data_as(fp), b, reg.ctypes.data_as(fp), cls.ctypes.data_as(fp)))
        return reg, cls

    def postprocess(self, regressors, classificators, padding=(0.0, 0.0, 0.0, 0.0), max_detections: int = 64, trace: dict | None = None):
        """decode -> sigmoid -> threshold -> weighted NMS -> letterbox removal on raw tensors of ONE frame
        ([N,16], [N,1]) or a batch ([B,N,16], [B,N,1]); returns list[Detection] (or a list per frame)."""
        reg = np.ascontiguousarray(regressors, np.float32)
        cls = np.ascontiguousarray(classificators, np.float32)
        single = reg.ndim == 2
        reg = reg.reshape(-1, self.num_anchors, 16)
        cls = cls.reshape(-1, self.num_anchors, 1)
        b = reg.shape[0]
        pad = np.ascontiguousarray(np.broadcast_to(np.asarray(padding, np.float64).reshape(-1, 4), (b, 4)))
        out = (CDetection * (max_detections * b))()
        n = (C.c_int * b)()
        cap_s = self.num_anchors
        sa = np.full((b, cap_s), -1, np.int32)
        sc = np.full((b, cap_s), -1, np.int32)
        ns = (C.c_int * b)()
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        check(lib().fdl_detector_postprocess(self._h, reg.ctypes.data_as(fp), cls.ctypes.data_as(fp), b, pad.ctypes.data_as(C.POINTER(C.c_double)),
                                             out, max_detections, n, sa.ctypes.data_as(ip), sc.ctypes.data_as(ip), cap_s, ns))
        res = [[Detection._from(out[i * max_detections + k]) for k in range(n[i])] for i in range(b)]
        if trace is not None:
            trace["survivors"] = [sa[i, :ns[i]].tolist() for i in range(b)]
            trace["survivor_cluster"] = [sc[i, :ns[i]].tolist() for i in range(b)]
        return res[0] if single else res


class FaceLandmark:
    """FaceMesh-468.  ``FaceLandmark::new`` / ``infer`` (face_landmark.rs:208, :232)."""

    def __init__(self, model_path: str | None = None, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fdl_landmark_create(os.fsencode(model_path) if model_path else None, device, C.byref(self._h)))
        self.device = device
        self.net = _Net(lib().fdl_landmark_net(self._h))
        self.last_face_flag = None

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_landmark_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def infer(self, image, roi: Rect | None = None):
        img, _keep = _image(image)
        out = (CLandmark * NUM_LANDMARKS)()
        n = C.c_int()
        flag = C.c_float()
        croi = roi._c() if roi is not None else None
        check(lib().fdl_landmark_infer(self._h, C.byref(img), C.byref(croi) if croi is not None else None, out, C.byref(n), C.byref(flag)))
        self.last_face_flag = flag.value
        return _landmarks(out, n.value)

    def forward(self, tensor: np.ndarray):
        x = np.ascontiguousarray(tensor, np.float32)
        b = x.shape[0]
        lm = np.empty((b, 1404), np.float32)
        fl = np.empty((b, 1), np.float32)
        fp = C.POINTER(C.c_float)
        check(lib().fdl_landmark_forward(self._h, x.ctypes.data_as(fp), b, lm.ctypes.data_as(fp), fl.ctypes.data_as(fp)))
        return lm, fl


class IrisLandmark:
    """Iris model.  ``IrisLandmark::new`` / ``infer`` (iris_landmark.rs:142, :158)."""

    def __init__(self, model_path: str | None = None, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fdl_iris_create(os.fsencode(model_path) if model_path else None, device, C.byref(self._h)))
        self.device = device
        self.net = _Net(lib().fdl_iris_net(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_iris_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def infer(self, image, roi: Rect | None = None, is_right_eye: bool | None = None) -> IrisResults:
        img, _keep = _image(image)
        contour = (CLandmark * _lib.NUM_EYE_CONTOUR)()
        iris = (CLandmark * _lib.NUM_IRIS)()
        croi = roi._c() if roi is not None else None
        check(lib().fdl_iris_infer(self._h, C.byref(img), C.byref(croi) if croi is not None else None, 1 if is_right_eye else 0, contour, iris))
        return IrisResults(_landmarks(contour, _lib.NUM_EYE_CONTOUR), _landmarks(iris, _lib.NUM_IRIS))

    def forward(self, tensor: np.ndarray):
        x = np.ascontiguousarray(tensor, np.float32)
        b = x.shape[0]
        eye = np.empty((b, 213), np.float32)
        ir = np.empty((b, 15), np.float32)
        fp = C.POINTER(C.c_float)
        check(lib().fdl_iris_forward(self._h, x.ctypes.data_as(fp), b, eye.ctypes.data_as(fp), ir.ctypes.data_as(fp)))
        return eye, ir


# ------------------------------------------------------------------------------------------------
def face_detection_to_roi(face_detection: Detection, image_size, size_mode: SizeMode | None = None, device: int = 0) -> Rect:
    """face_landmark.rs:180-198.  ``image_size`` = (width, height)."""
    out = CRect()
    det = face_detection._c()
    check(lib().fdl_face_detection_to_roi(device, C.byref(det), int(image_size[0]), int(image_size[1]),
                                          -1 if size_mode is None else int(size_mode), C.byref(out)))
    return Rect._from(out)


def iris_roi_from_face_landmarks(face_landmarks, image_size, device: int = 0):
    """iris_landmark.rs:268-292 -> (left_eye_roi, right_eye_roi)."""
    n = len(face_landmarks)
    arr = (CLandmark * max(n, 1))()
    for i, l in enumerate(face_landmarks):
        arr[i].x, arr[i].y, arr[i].z = l.x, l.y, l.z
    left, right = CRect(), CRect()
    check(lib().fdl_iris_roi_from_face_landmarks(device, arr, n, int(image_size[0]), int(image_size[1]), C.byref(left), C.byref(right)))
    return Rect._from(left), Rect._from(right)


def _landmark_array(points):
    """list[Landmark] / [n,3] array -> (CLandmark * n)."""
    n = len(points)
    arr = (CLandmark * max(n, 1))()
    for i, l in enumerate(points):
        if hasattr(l, "x"):
            arr[i].x, arr[i].y, arr[i].z = l.x, l.y, l.z
        else:
            arr[i].x, arr[i].y, arr[i].z = float(l[0]), float(l[1]), float(l[2])
    return arr, n


def eye_to_face_landmark_index(is_right_eye: bool) -> np.ndarray:
    """LEFT_/RIGHT_EYE_TO_FACE_LANDMARK_INDEX (iris_landmark.rs:64-95) as the library holds them."""
    out = (C.c_int32 * _lib.NUM_EYE_CONTOUR)()
    check(lib().fdl_eye_to_face_landmark_index(1 if is_right_eye else 0, out))
    return np.array(out[:], np.int32)


def update_face_landmarks_with_iris_results(face_landmarks, iris_data_left, iris_data_right, device: int = 0):
    """iris_landmark.rs:380-398.  ``iris_data_*``: IrisResults (or a [<=71,3] contour array).  -> list[Landmark] (468)."""
    face, n = _landmark_array(face_landmarks)
    lc, nl = _landmark_array(iris_data_left.contour if isinstance(iris_data_left, IrisResults) else iris_data_left)
    rc_, nr = _landmark_array(iris_data_right.contour if isinstance(iris_data_right, IrisResults) else iris_data_right)
    out = (CLandmark * _lib.NUM_FACE_LANDMARKS)()
    check(lib().fdl_update_face_landmarks_with_iris_results(device, face, n, lc, nl, rc_, nr, out))
    return _landmarks(out, _lib.NUM_FACE_LANDMARKS)


def get_iris_diameter(iris_landmarks, image_size, device: int = 0) -> float:
    """iris_landmark.rs:401-418 (private in the reference): iris diameter in pixels."""
    arr, n = _landmark_array(iris_landmarks)
    out = C.c_double()
    check(lib().fdl_iris_diameter(device, arr, n, int(image_size[0]), int(image_size[1]), C.byref(out)))
    return out.value


def get_iris_depth(iris_landmarks, focal_length_mm: float, iris_size_px: float, image_size, device: int = 0) -> float:
    """iris_landmark.rs:421-433 (private in the reference): iris distance in millimetres."""
    arr, n = _landmark_array(iris_landmarks)
    out = C.c_double()
    check(lib().fdl_iris_depth(device, arr, n, float(focal_length_mm), float(iris_size_px), int(image_size[0]), int(image_size[1]), C.byref(out)))
    return out.value


def image_to_tensor(image, roi: Rect | None, output_size, keep_aspect_ratio: bool, output_range=(0.0, 1.0), flip_horizontal: bool = False,
                    device: int = 0):
    """transform.rs:188-309 (private in the reference; exposed for parity checks).
    Returns (tensor f32 [h,w,3], padding (l,t,r,b), uint8 image before normalisation)."""
    img, _keep = _image(image)
    w, h = int(output_size[0]), int(output_size[1])
    tensor = np.empty((h, w, 3), np.float32)
    u8 = np.empty((h, w, 3), np.uint8)
    pad = (C.c_double * 4)()
    croi = roi._c() if roi is not None else None
    check(lib().fdl_image_to_tensor(device, C.byref(img), C.byref(croi) if croi is not None else None, w, h, 1 if keep_aspect_ratio else 0,
                                    float(output_range[0]), float(output_range[1]), 1 if flip_horizontal else 0,
                                    tensor.ctypes.data_as(C.POINTER(C.c_float)), u8.ctypes.data_as(C.POINTER(C.c_uint8)), pad))
    return tensor, tuple(pad), u8


def project_landmarks(data, tensor_size, image_size, padding, roi: Rect | None, flip_horizontal: bool, device: int = 0):
    """transform.rs:351-432 (private in the reference; exposed for parity checks) -> [K,3] float64."""
    raw = np.ascontiguousarray(data, np.float32).reshape(-1)
    n = raw.size // 3
    out = (CLandmark * n)()
    pad = (C.c_double * 4)(*[float(p) for p in padding])
    croi = roi._c() if roi is not None else None
    check(lib().fdl_project_landmarks(device, raw.ctypes.data_as(C.POINTER(C.c_float)), n, int(tensor_size[0]), int(tensor_size[1]),
                                      int(image_size[0]), int(image_size[1]), pad, C.byref(croi) if croi is not None else None,
                                      1 if flip_horizontal else 0, out))
    return np.array([[out[i].x, out[i].y, out[i].z] for i in range(n)], np.float64)


# ------------------------------------------------------------------------------------------------
@dataclass
class FaceResult:
    roi: Rect
    face_flag_logit: float
    landmarks: np.ndarray | None          # [468,3] float32 or None when the face flag gate rejected the crop
    left_eye_roi: Rect | None
    right_eye_roi: Rect | None
    left_contour: np.ndarray | None       # [71,3]
    left_iris: np.ndarray | None          # [5,3]
    right_contour: np.ndarray | None
    right_iris: np.ndarray | None
    refined_landmarks: np.ndarray | None = None   # [468,3]: update_face_landmarks_with_iris_results (Pipeline(refine_landmarks=True))
    iris_diameter_px: tuple | None = None         # (left, right): get_iris_diameter
    iris_depth_mm: tuple | None = None            # (left, right): get_iris_depth (Pipeline(focal_length_mm=...))


@dataclass
class FrameResult:
    detections: list
    faces: list
    n_total_detections: int = 0    # what weighted NMS produced; > len(detections) when the record was truncated at 32


class Pipeline:
    """Batched detect -> landmark -> iris, the call sequence of lib.rs:20-40 for a batch of frames."""

    def __init__(self, detector_model: FaceDetectionModel = FaceDetectionModel.BackCamera, frame_size=(1920, 1080), max_batch: int = 64,
                 max_faces: int = 1, run_landmarks: bool = True, run_iris: bool = True, model_dir: str | None = None, device: int = 0,
                 zero_copy_host: bool = False, refine_landmarks: bool = False, focal_length_mm: float = 0.0, allow_truncated: bool = False):
        self._h = C.c_void_p()
        self.allow_truncated = bool(allow_truncated)   # collect() raises FdlError(FDL_ERR_CAPACITY) for > 32 detections in a frame unless set
        self._dir = os.fsencode(model_dir) if model_dir else None
        cfg = CPipelineConfig(int(detector_model), device, max_batch, max_faces, int(frame_size[0]), int(frame_size[1]),
                              1 if run_landmarks else 0, 1 if (run_iris and run_landmarks) else 0, self._dir, 1 if zero_copy_host else 0,
                              1 if refine_landmarks else 0, float(focal_length_mm))
        self.refine_landmarks, self.focal_length_mm = bool(refine_landmarks), float(focal_length_mm)
        check(lib().fdl_pipeline_create(C.byref(cfg), C.byref(self._h)))
        self.max_batch, self.max_faces = max_batch, max_faces
        self.frame_size = (int(frame_size[0]), int(frame_size[1]))
        self.run_landmarks, self.run_iris = run_landmarks, run_iris and run_landmarks
        self.device = device
        self._frames = (CFrameResult * max_batch)()
        self._faces = (CFaceResult * (max_batch * max_faces))()
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _images(self, frames):
        """frames: list of [H,W,3] arrays/tensors, or one [B,H,W,3] numpy array / torch tensor."""
        keep = frames
        if hasattr(frames, "shape") and len(frames.shape) == 4:
            b, h, w, _ = frames.shape
            if isinstance(frames, np.ndarray):
                frames = np.ascontiguousarray(frames)
                base, mem = frames.ctypes.data, _lib.MEM_HOST
            else:
                frames = frames.contiguous()
                base, mem = frames.data_ptr(), (_lib.MEM_DEVICE if frames.is_cuda else _lib.MEM_HOST)
            keep = frames
            arr = (CImage * b)()
            for i in range(b):
                arr[i] = CImage(base + i * h * w * 3, w, h, w * 3, mem, 0)
            return arr, b, keep
        imgs = [_image(f) for f in frames]
        arr = (CImage * len(imgs))(*[i[0] for i in imgs])
        return arr, len(imgs), [i[1] for i in imgs]

    def submit(self, frames) -> int:
        arr, n, keep = self._images(frames)
        t = C.c_int()
        check(lib().fdl_pipeline_submit(self._h, arr, n, C.byref(t)))
        self._keep[t.value] = keep
        return t.value

    def submit_jpeg(self, files) -> int:
        """``files``: JPEG byte strings (bytes / bytearray / uint8 arrays), or one ``(uint8 buffer, offsets, lengths)`` triple for files
        that already sit in one (ideally pinned) arena -- the form that lets the copy engine read them in place."""
        ptrs, lens, n, keep = _jpeg_args(files)
        t = C.c_int()
        check(lib().fdl_pipeline_submit_jpeg(self._h, ptrs, lens, n, C.byref(t)))
        self._keep[t.value] = keep
        return t.value

    def run_jpeg(self, files):
        return self.collect(self.submit_jpeg(files))

    def collect_raw(self, ticket: int) -> int:
        """Waits for `ticket`; results stay in the ctypes arrays (self._frames / self._faces). Returns n."""
        n = C.c_int()
        rc = lib().fdl_pipeline_collect(self._h, ticket, self._frames, self._faces, C.byref(n))
        self._keep.pop(ticket, None)
        if rc == _lib.FDL_ERR_CAPACITY and self.allow_truncated:
            return n.value          # the records are delivered; FrameResult.n_total_detections tells which frames overflowed
        check(rc)
        return n.value

    def collect(self, ticket: int):
        n = self.collect_raw(ticket)
        return [self._frame(i) for i in range(n)]

    def run(self, frames):
        return self.collect(self.submit(frames))

    def _frame(self, i) -> FrameResult:
        fr = self._frames[i]
        dets = [Detection._from(fr.detections[k]) for k in range(fr.n_detections)]
        faces = []
        if self.run_landmarks:
            for f in range(fr.n_faces):
                c = self._faces[i * self.max_faces + f]
                has = bool(c.has_landmarks)
                lm = np.array(c.landmarks[:], np.float32).reshape(-1, 3) if has else None
                iris_ok = has and self.run_iris
                g = lambda a: np.array(a[:], np.float32).reshape(-1, 3)
                faces.append(FaceResult(Rect._from(c.face_roi), float(c.face_flag_logit), lm,
                                        Rect._from(c.eye_roi[0]) if has else None, Rect._from(c.eye_roi[1]) if has else None,
                                        g(c.eye_contour[0]) if iris_ok else None, g(c.iris[0]) if iris_ok else None,
                                        g(c.eye_contour[1]) if iris_ok else None, g(c.iris[1]) if iris_ok else None,
                                        g(c.refined_landmarks) if (iris_ok and self.refine_landmarks) else None,
                                        (c.iris_diameter_px[0], c.iris_diameter_px[1]) if iris_ok else None,
                                        (c.iris_depth_mm[0], c.iris_depth_mm[1]) if (iris_ok and self.focal_length_mm > 0) else None))
        return FrameResult(dets, faces, int(fr.n_total_detections))

    @property
    def last_device_ms(self) -> float:
        return lib().fdl_pipeline_last_device_ms(self._h)

    @property
    def stage_ms(self):
        out = (C.c_float * 10)()
        check(lib().fdl_pipeline_stage_ms(self._h, out))
        return list(out)


# ------------------------------------------------------------------------------------------------
# frame ingest: utils.rs:8-21 convert_image_to_mat (imdecode + BGR2RGB), decoded on the device
def _buffer_address(b):
    """(address, length, keepalive) of bytes / bytearray / numpy uint8 / torch uint8 (host)."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, np.uint8)
        return a.ctypes.data, a.size, a
    if hasattr(b, "data_ptr"):
        return b.data_ptr(), b.numel(), b
    a = np.frombuffer(b, np.uint8)
    return a.ctypes.data, a.size, (a, b)


def _jpeg_args(files):
    if isinstance(files, tuple) and len(files) == 3 and not isinstance(files[0], (bytes, bytearray)):
        base, blen, keep = _buffer_address(files[0])
        offs, lens_ = np.asarray(files[1], np.int64), np.asarray(files[2], np.int64)
        if len(offs) != len(lens_) or (len(offs) and (offs.min() < 0 or (offs + lens_).max() > blen)):
            raise FdlError(_lib.FDL_ERR_INVALID, "JPEG arena: offsets / lengths out of range")
        n = len(offs)
        ptrs = (C.c_void_p * max(n, 1))(*[base + int(o) for o in offs])
        lens = (C.c_size_t * max(n, 1))(*[int(x) for x in lens_])
        return ptrs, lens, n, keep
    keeps, n = [], len(files)
    ptrs, lens = (C.c_void_p * max(n, 1))(), (C.c_size_t * max(n, 1))()
    for i, f in enumerate(files):
        a, ln, k = _buffer_address(f)
        ptrs[i], lens[i] = a, ln
        keeps.append(k)
    return ptrs, lens, n, keeps


def jpeg_info(data):
    """(width, height, components) from the header alone (host; no GPU needed)."""
    a, ln, _k = _buffer_address(data)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().fdl_jpeg_info(a, ln, C.byref(w), C.byref(h), C.byref(c)))
    return w.value, h.value, c.value


def convert_image_to_mat(im_bytes, device: int = 0) -> np.ndarray:
    """utils.rs:8-21: JPEG bytes -> RGB uint8 [H,W,3] (imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB)), decoded on the device."""
    a, ln, _k = _buffer_address(im_bytes)
    w, h = C.c_int(), C.c_int()
    check(lib().fdl_jpeg_info(a, ln, C.byref(w), C.byref(h), None))
    out = np.empty((h.value, w.value, 3), np.uint8)
    check(lib().fdl_decode_jpeg(device, a, ln, out.ctypes.data, out.size, C.byref(w), C.byref(h)))
    return out


class JpegDecoder:
    """Batched device decoder (fdl_jpeg_decode): ``decode(files)`` -> list of RGB uint8 arrays; ``decode_to_device(files)`` -> one
    CUDA uint8 tensor holding the images back to back + (offsets, widths, heights)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().fdl_jpeg_decoder_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_jpeg_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sizes(self, ptrs, lens, n):
        offs, ws, hs = (C.c_int64 * n)(), (C.c_int32 * n)(), (C.c_int32 * n)()
        rc = lib().fdl_jpeg_decode(self._h, ptrs, lens, n, None, 0, _lib.MEM_HOST, offs, ws, hs)
        if rc != _lib.FDL_ERR_CAPACITY:
            check(rc)
        total = (int(offs[n - 1]) + int(ws[n - 1]) * int(hs[n - 1]) * 3 + 3) & ~3
        return offs, ws, hs, total

    def decode(self, files):
        ptrs, lens, n, _keep = _jpeg_args(files)
        offs, ws, hs, total = self._sizes(ptrs, lens, n)
        out = np.empty(total, np.uint8)
        check(lib().fdl_jpeg_decode(self._h, ptrs, lens, n, out.ctypes.data, out.size, _lib.MEM_HOST, offs, ws, hs))
        return [out[offs[i]:offs[i] + ws[i] * hs[i] * 3].reshape(hs[i], ws[i], 3) for i in range(n)]

    def decode_to_device(self, files):
        import torch
        ptrs, lens, n, _keep = _jpeg_args(files)
        offs, ws, hs, total = self._sizes(ptrs, lens, n)
        out = torch.empty(total, dtype=torch.uint8, device="cuda:%d" % self.device)
        check(lib().fdl_jpeg_decode(self._h, ptrs, lens, n, out.data_ptr(), total, _lib.MEM_DEVICE, offs, ws, hs))
        return out, list(offs), list(ws), list(hs)


# ------------------------------------------------------------------------------------------------
# drawing: render.rs (Color / Colors / Annotation, detections_to_render_data, landmarks_to_render_data, render_to_image) and the
# connection tables of face_landmark.rs:35-160 / iris_landmark.rs:44-60.  The bookkeeping is host side; the pixels are painted on the device.
@dataclass
class Color:  # render.rs:7-27
    r: int = 0
    g: int = 0
    b: int = 0
    a: int | None = None


class Colors:  # render.rs:29-68 (the ones lib.rs uses and the primaries)
    BLACK = Color(0, 0, 0)
    RED = Color(255, 0, 0)
    GREEN = Color(0, 255, 0)
    BLUE = Color(0, 0, 255)
    PINK = Color(255, 0, 255)
    WHITE = Color(255, 255, 255)


@dataclass
class Annotation:  # render.rs:208-213: data = [("point", x, y) | ("line", x0, y0, x1, y1) | ("rect", l, t, r, b) | ("filled_rect", l, t, r, b)]
    data: list
    normalized_positions: bool
    thickness: float
    color: Color


FACE_LANDMARK_CONNECTIONS = [
    (61, 146), (146, 91), (91, 181), (181, 84), (84, 17), (17, 314), (314, 405), (405, 321), (321, 375), (375, 291), (61, 185), (185, 40), (40, 39),
    (39, 37), (37, 0), (0, 267), (267, 269), (269, 270), (270, 409), (409, 291), (78, 95), (95, 88), (88, 178), (178, 87), (87, 14), (14, 317),
    (317, 402), (402, 318), (318, 324), (324, 308), (78, 191), (191, 80), (80, 81), (81, 82), (82, 13), (13, 312), (312, 311), (311, 310), (310, 415),
    (415, 308), (33, 7), (7, 163), (163, 144), (144, 145), (145, 153), (153, 154), (154, 155), (155, 133), (33, 246), (246, 161), (161, 160),
    (160, 159), (159, 158), (158, 157), (157, 173), (173, 133), (46, 53), (53, 52), (52, 65), (65, 55), (70, 63), (63, 105), (105, 66), (66, 107),
    (263, 249), (249, 390), (390, 373), (373, 374), (374, 380), (380, 381), (381, 382), (382, 362), (263, 466), (466, 388), (388, 387), (387, 386),
    (386, 385), (385, 384), (384, 398), (398, 362), (276, 283), (283, 282), (282, 295), (295, 285), (300, 293), (293, 334), (334, 296), (296, 336),
    (10, 338), (338, 297), (297, 332), (332, 284), (284, 251), (251, 389), (389, 356), (356, 454), (454, 323), (323, 361), (361, 288), (288, 397),
    (397, 365), (365, 379), (379, 378), (378, 400), (400, 377), (377, 152), (152, 148), (148, 176), (176, 149), (149, 150), (150, 136), (136, 172),
    (172, 58), (58, 132), (132, 93), (93, 234), (234, 127), (127, 162), (162, 21), (21, 54), (54, 103), (103, 67), (67, 109), (109, 10)]
EYE_LANDMARK_CONNECTIONS = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (9, 10), (10, 11), (11, 12), (12, 13), (13, 14), (0, 9), (8, 14)]
MAX_EYE_LANDMARK = len(EYE_LANDMARK_CONNECTIONS)


def detections_to_render_data(detections, bounds_color=None, keypoint_color=None, line_width: int = 1, point_width: int = 3,
                              normalized_positions: bool = True, output=None):
    """render.rs:262-313: one annotation with every detection's bounds (if bounds_color and line_width > 0), one with every row of
    every detection as a point (if keypoint_color and point_width > 0)."""
    out = list(output) if output is not None else []
    if bounds_color is not None and line_width > 0:
        rects = []
        for d in detections:
            b = d.bbox()
            rects.append(("rect", b.xmin, b.ymin, b.xmax, b.ymax))
        out.append(Annotation(rects, normalized_positions, float(line_width), bounds_color))
    if keypoint_color is not None and point_width > 0:
        pts = [("point", float(row[0]), float(row[1])) for d in detections for row in np.asarray(d.data).reshape(-1, 2)]
        out.append(Annotation(pts, normalized_positions, float(point_width), keypoint_color))
    return out


def landmarks_to_render_data(landmarks, landmark_connections, landmark_color=None, connection_color=None, thickness=None,
                             normalized_positions=None, output=None):
    """render.rs:315-359: the connection lines, then the points."""
    lc = landmark_color if landmark_color is not None else Colors.RED
    cc = connection_color if connection_color is not None else Colors.RED
    th = float(np.float32(thickness)) if thickness is not None else 1.0
    norm = True if normalized_positions is None else bool(normalized_positions)
    xy = [(l.x, l.y) if hasattr(l, "x") else (float(l[0]), float(l[1])) for l in landmarks]
    lines = [("line", xy[a][0], xy[a][1], xy[b][0], xy[b][1]) for a, b in landmark_connections]
    points = [("point", x, y) for x, y in xy]
    out = list(output) if output is not None else []
    out += [Annotation(lines, norm, th, cc), Annotation(points, norm, th, lc)]
    return out


def face_landmarks_to_render_data(face_landmarks, landmark_color, connection_color, thickness=None, output=None):
    """face_landmark.rs:324-340."""
    return landmarks_to_render_data(face_landmarks, FACE_LANDMARK_CONNECTIONS, landmark_color, connection_color, 2.0 if thickness is None else thickness,
                                    True, output)


def eye_landmarks_to_render_data(eye_contour, landmark_color, connection_color, thickness=None, output=None):
    """iris_landmark.rs:312-328: the first 15 contour points and their connections."""
    return landmarks_to_render_data(list(eye_contour)[:MAX_EYE_LANDMARK], EYE_LANDMARK_CONNECTIONS, landmark_color, connection_color,
                                    2.0 if thickness is None else thickness, True, output)


def iris_landmarks_to_render_data(iris_landmarks, landmark_color=None, oval_color=None, thickness=None, image_size=None, output=None, device: int = 0):
    """iris_landmark.rs:330-376: the iris circle as an "oval" (which render.rs:447-462 draws as a hollow rectangle) and the 5 points."""
    w, h = image_size if image_size is not None else (-1, -1)
    th = 1.0 if thickness is None else float(thickness)
    ann = []
    if oval_color is not None:
        if w < 2 or h < 2:
            raise FdlError(_lib.FDL_ERR_INVALID, "oval_color requires a valid image_size arg")
        rad = get_iris_diameter(iris_landmarks, (w, h), device=device) / 2.0
        c = iris_landmarks[0]
        ann.append(Annotation([("rect", c.x - rad / w, c.y - rad / h, c.x + rad / w, c.y + rad / h)], True, th, oval_color))
    if landmark_color is not None:
        ann.append(Annotation([("point", l.x, l.y) for l in iris_landmarks], True, th, landmark_color))
    return (list(output) if output is not None else []) + ann


_PRIM_KIND = {"point": 0, "line": 1, "rect": 2, "filled_rect": 3}


def render_to_image(annotations, image, blend_mode=None, device: int = 0) -> np.ndarray:
    """render.rs:361-479 -> RGBA uint8 [H,W,4] (the reference's DynamicImage::ImageRgba8); painted on the device."""
    img, _keep = _image(image)
    items = [(a, it) for a in annotations for it in a.data]
    prims = (_lib.CPrimitive * max(len(items), 1))()
    for k, (a, it) in enumerate(items):
        p = prims[k]
        p.kind, p.normalized, p.thickness = _PRIM_KIND[it[0]], 1 if a.normalized_positions else 0, float(a.thickness)
        vals = list(it[1:5]) + [0.0, 0.0]
        p.a, p.b, p.c, p.d = (float(v) for v in vals[:4])
        col = it[5] if it[0] == "filled_rect" and len(it) > 5 else a.color        # FilledRectOrOval carries its own fill (render.rs:131-135)
        p.r, p.g, p.b_, p.alpha = col.r & 255, col.g & 255, col.b & 255, (255 if col.a is None else col.a) & 255
    out = np.empty((img.height, img.width, 4), np.uint8)
    check(lib().fdl_render_to_image(device, C.byref(img), prims, len(items), out.ctypes.data, out.size, _lib.MEM_HOST))
    return out


class Pool(Pipeline):
    """Every GPU of a box behind one handle (fdl_pool): one pipeline + one host worker thread per device, least-loaded dispatch,
    pool-wide tickets.  Same ``submit`` / ``submit_jpeg`` / ``collect`` / ``run`` as ``Pipeline``; up to ``depth`` tickets in flight."""

    def __init__(self, devices, detector_model: FaceDetectionModel = FaceDetectionModel.BackCamera, frame_size=(1920, 1080), max_batch: int = 64,
                 max_faces: int = 1, run_landmarks: bool = True, run_iris: bool = True, model_dir: str | None = None,
                 zero_copy_host: bool = False, refine_landmarks: bool = False, focal_length_mm: float = 0.0, allow_truncated: bool = False):
        self._h = C.c_void_p()
        self.allow_truncated = bool(allow_truncated)
        self._dir = os.fsencode(model_dir) if model_dir else None
        cfg = CPipelineConfig(int(detector_model), 0, max_batch, max_faces, int(frame_size[0]), int(frame_size[1]),
                              1 if run_landmarks else 0, 1 if (run_iris and run_landmarks) else 0, self._dir, 1 if zero_copy_host else 0,
                              1 if refine_landmarks else 0, float(focal_length_mm))
        self.refine_landmarks, self.focal_length_mm = bool(refine_landmarks), float(focal_length_mm)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        check(lib().fdl_pool_create(C.byref(cfg), devs, len(devices), C.byref(self._h)))
        self.devices = list(devices)
        self.max_batch, self.max_faces = max_batch, max_faces
        self.frame_size = (int(frame_size[0]), int(frame_size[1]))
        self.run_landmarks, self.run_iris = run_landmarks, run_iris and run_landmarks
        self._frames = (CFrameResult * max_batch)()
        self._faces = (CFaceResult * (max_batch * max_faces))()
        self._keep = {}
        self.last_device_index = None

    @property
    def depth(self) -> int:
        return lib().fdl_pool_depth(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().fdl_pool_destroy(self._h)
            self._h = None

    def submit(self, frames) -> int:
        arr, n, keep = self._images(frames)
        t = C.c_int()
        check(lib().fdl_pool_submit(self._h, arr, n, C.byref(t)))
        self._keep[t.value] = (keep, arr)
        return t.value

    def submit_jpeg(self, files) -> int:
        ptrs, lens, n, keep = _jpeg_args(files)
        t = C.c_int()
        check(lib().fdl_pool_submit_jpeg(self._h, ptrs, lens, n, C.byref(t)))
        self._keep[t.value] = (keep, ptrs, lens)
        return t.value

    def collect_raw(self, ticket: int) -> int:
        n, dev = C.c_int(), C.c_int()
        rc = lib().fdl_pool_collect(self._h, ticket, self._frames, self._faces, C.byref(n), C.byref(dev))
        self._keep.pop(ticket, None)
        self.last_device_index = dev.value
        if rc == _lib.FDL_ERR_CAPACITY and self.allow_truncated:
            return n.value
        check(rc)
        return n.value

    @property
    def last_device_ms(self):
        raise AttributeError("per-device timings are not exposed by the pool")

    @property
    def stage_ms(self):
        raise AttributeError("per-device timings are not exposed by the pool")


def letterbox_row_plan(frame_size, input_size: int):
    """Zero-copy ingest plan of the detector letterbox: ``(row_pos[H], info)`` or ``None`` when the rows are not gathered."""
    w, h = frame_size
    rp = (C.c_int32 * h)()
    info = (C.c_int32 * 4)()
    rc = lib().fdl_letterbox_row_plan(w, h, input_size, rp, info)
    if rc < 0:
        check(rc)
    if rc == 0:
        return None
    return np.array(rp[:], np.int32), {"rows_per_frame": info[0], "period_src_rows": info[1], "periods_per_frame": info[2], "copies": info[3]}


def device_count() -> int:
    return lib().fdl_device_count()


def launch_count() -> int:
    return int(lib().fdl_launch_count())
