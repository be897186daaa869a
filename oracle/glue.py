"""numpy restatement of the reference's Rust glue code -- oracle side.

Test infrastructure; see ``oracle/__init__.py``.  Every function cites the
reference lines it follows (paths relative to
/root/reference/src/face_detection_lite/).  Precision follows the Rust types:
``Detection.data`` is f32, ``BBox``/``Rect``/``Landmark`` are f64, and every
f32<->f64 cast in the reference is reproduced.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import cv2
import numpy as np

from . import cv_ops

f32 = np.float32

# face_detection.rs:133-139
RAW_SCORE_LIMIT = f32(80.0)
MIN_SCORE = f32(0.5)
MIN_SUPPRESSION_THRESHOLD = f32(0.3)

# face_detection.rs:117-123
FRONT_CAMERA, BACK_CAMERA, SHORT, FULL, FULL_SPARSE = 0, 1, 2, 3, 4
MODEL_FILES = {  # face_detection.rs:125-129
    FRONT_CAMERA: "face_detection_front.tflite",
    BACK_CAMERA: "face_detection_back.tflite",
    SHORT: "face_detection_short_range.tflite",
    FULL: "face_detection_full_range.tflite",
    FULL_SPARSE: "face_detection_full_range_sparse.tflite",
}
# SSDOptions, face_detection.rs:39-85: (num_layers, input size, strides, interpolated_scale_aspect_ratio)
SSD_OPTIONS = {
    FRONT_CAMERA: (4, 128, [8, 16, 16, 16], 1.0),
    BACK_CAMERA: (4, 256, [16, 32, 32, 32], 1.0),
    SHORT: (4, 128, [8, 16, 16, 16], 1.0),
    FULL: (1, 192, [4, 0, 0, 0], 0.0),
    FULL_SPARSE: (1, 192, [4, 0, 0, 0], 0.0),
}
SIZE_DEFAULT, SQUARE_LONG, SQUARE_SHORT = 0, 1, 2  # transform.rs:15-24


# ----------------------------------------------------------------------------
# types.rs
# ----------------------------------------------------------------------------
@dataclass
class Rect:  # types.rs:24-97
    x_center: float
    y_center: float
    width: float
    height: float
    rotation: float
    normalized: bool

    def size(self):  # :52-59
        if self.normalized:
            return (self.width, self.height)
        return (float(int(self.width)), float(int(self.height)))

    def scaled(self, size, normalize):  # :62-77
        if self.normalized == normalize:
            return Rect(self.x_center, self.y_center, self.width, self.height, self.rotation, self.normalized)
        sx, sy = (1.0 / size[0], 1.0 / size[1]) if normalize else size
        return Rect(self.x_center * sx, self.y_center * sy, self.width * sx, self.height * sy,
                    self.rotation, normalize)

    def points(self):  # :80-96  (TL, TR, BR, BL; rotation about the centre)
        x, y = self.x_center, self.y_center
        w, h = self.width / 2.0, self.height / 2.0
        pts = [(x - w, y - h), (x + w, y - h), (x + w, y + h), (x - w, y + h)]
        if self.rotation != 0.0:
            s, c = math.sin(self.rotation), math.cos(self.rotation)
            pts = [(x + (px - x) * c - (py - y) * s, y + (px - x) * s + (py - y) * c) for px, py in pts]
        return pts


@dataclass
class BBox:  # types.rs:99-174
    xmin: float
    ymin: float
    xmax: float
    ymax: float

    def width(self): return self.xmax - self.xmin
    def height(self): return self.ymax - self.ymin
    def empty(self): return self.width() <= 0.0 or self.height() <= 0.0
    def normalized(self): return self.xmin >= -1.0 and self.xmax < 2.0 and self.ymin >= -1.0  # :134-136 (sic)
    def area(self): return 0.0 if self.empty() else self.width() * self.height()

    def intersect(self, o):  # :139-150
        xmin, ymin = max(self.xmin, o.xmin), max(self.ymin, o.ymin)
        xmax, ymax = min(self.xmax, o.xmax), min(self.ymax, o.ymax)
        if xmin < xmax and ymin < ymax:
            return BBox(xmin, ymin, xmax, ymax)
        return None

    def absolute(self, size):  # :168-173
        if not self.normalized():
            return self
        return BBox(self.xmin * size[0], self.ymin * size[1], self.xmax * size[0], self.ymax * size[1])


@dataclass
class Detection:  # types.rs:189-246
    data: np.ndarray  # [8,2] f32
    score: np.float32
    anchor: int = -1  # not in the reference: index of the cluster's top anchor (for parity reports)

    def bbox(self):  # :215-221
        d = self.data
        return BBox(float(d[0, 0]), float(d[0, 1]), float(d[1, 0]), float(d[1, 1]))

    def keypoint(self, k):  # :209-212
        return (self.data[k + 2, 0], self.data[k + 2, 1])


@dataclass
class Landmark:  # types.rs:176-187
    x: float
    y: float
    z: float


# ----------------------------------------------------------------------------
# face_detection.rs
# ----------------------------------------------------------------------------
def ssd_generate_anchors(model_type: int) -> np.ndarray:
    """face_detection.rs:366-413.  f32 arithmetic throughout."""
    num_layers, size, strides, interp = SSD_OPTIONS[model_type]
    anchors = []
    layer_id = 0
    while layer_id < num_layers:
        last = layer_id
        repeats = 0
        while last < num_layers and strides[last] == strides[layer_id]:
            last += 1
            repeats += 2 if interp == 1.0 else 1
        stride = strides[layer_id]
        fm_h = size // stride
        fm_w = size // stride
        for y in range(fm_h):
            yc = (f32(y) + f32(0.5)) / f32(fm_h)
            for x in range(fm_w):
                xc = (f32(x) + f32(0.5)) / f32(fm_w)
                for _ in range(repeats):
                    anchors.append((xc, yc))
        layer_id = last
    return np.array(anchors, dtype=np.float32).reshape(-1, 2)


def decode_boxes(raw_boxes: np.ndarray, anchors: np.ndarray, scale: float) -> np.ndarray:
    """face_detection.rs:269-296.  raw [1,N,16] f32 -> [N,8,2] f32."""
    raw = np.asarray(raw_boxes, np.float32).reshape(-1, 16)
    n = raw.shape[0]
    boxes = (raw / f32(scale)).reshape(n, 8, 2).astype(np.float32)
    boxes[:, 0, :] += anchors
    for i in range(2, 8):
        boxes[:, i, :] += anchors
    center = boxes[:, 0, :].copy()
    half = (boxes[:, 1, :] / f32(2.0)).astype(np.float32)
    boxes[:, 0, :] = center - half
    boxes[:, 1, :] = center + half
    return boxes


def sigmoid_f32(x):
    """transform.rs:111-113 in f32."""
    x = np.asarray(x, np.float32)
    # Rust's f32::exp is the host libm's expf (glibc: correctly rounded in practice).  numpy's float32 exp is a
    # SIMD approximation good to ~2 ulp, so evaluate in f64 and round once to f32 to model the correctly rounded value.
    e = np.exp(-x.astype(np.float64)).astype(np.float32)
    return (f32(1.0) / (f32(1.0) + e)).astype(np.float32)


def get_sigmoid_score(raw_scores: np.ndarray) -> np.ndarray:
    """face_detection.rs:300-314."""
    x = np.clip(np.asarray(raw_scores, np.float32), -RAW_SCORE_LIMIT, RAW_SCORE_LIMIT)
    return sigmoid_f32(x)


def convert_to_detections(boxes: np.ndarray, scores: np.ndarray):
    """face_detection.rs:317-362: score > 0.5 and xmax>xmin and ymax>ymin, anchor order."""
    scores = np.asarray(scores, np.float32).reshape(-1)
    dets = []
    for j in np.nonzero(scores > MIN_SCORE)[0]:
        b = boxes[j]
        if b[1, 0] > b[0, 0] and b[1, 1] > b[0, 1]:
            dets.append(Detection(b.copy(), scores[j], int(j)))
    return dets


# ----------------------------------------------------------------------------
# nms.rs
# ----------------------------------------------------------------------------
def overlap_similarity(b1: BBox, b2: BBox) -> float:
    """nms.rs:5-17 (f64 IoU)."""
    inter = b1.intersect(b2)
    if inter is None:
        return 0.0
    ia = inter.area()
    den = b1.area() + b2.area() - ia
    return ia / den if den > 0.0 else 0.0


def weighted_non_maximum_suppression(indexed_scores, detections, min_suppression_threshold, min_score,
                                     clusters=None):
    """nms.rs:56-124.  ``clusters`` (optional list) receives, per output, the anchor
    indices of the cluster members in candidate order (top first) -- the
    reference does not expose them; they define 'kept indices' for parity."""
    remaining_indexed = list(indexed_scores)
    outputs = []
    thr = float(f32(min_suppression_threshold))  # `min_suppression_threshold as f64`
    while remaining_indexed:
        det = detections[remaining_indexed[0][0]]
        if min_score is not None and det.score < min_score:
            break
        n_prev = len(remaining_indexed)
        det_bbox = det.bbox()
        remaining, candidates = [], []
        for index, score in remaining_indexed:
            sim = overlap_similarity(detections[index].bbox(), det_bbox)
            if sim > thr:
                candidates.append((index, score))
            else:
                remaining.append((index, score))
        weighted = Detection(det.data.copy(), det.score, det.anchor)
        if candidates:
            w = np.zeros((det.data.shape[0], 2), np.float32)
            total = f32(0.0)
            for index, score in candidates:
                total = f32(total + score)
                w = (w + detections[index].data * f32(score)).astype(np.float32)
            w = (w / total).astype(np.float32)
            weighted = Detection(w, det.score, det.anchor)
        outputs.append(weighted)
        if clusters is not None:
            clusters.append([detections[i].anchor for i, _ in candidates])
        if n_prev == len(remaining):
            break
        remaining_indexed = remaining
    return outputs


def non_maximum_suppression(detections, min_suppression_threshold=MIN_SUPPRESSION_THRESHOLD,
                            min_score=MIN_SCORE, weighted=True, clusters=None):
    """nms.rs:127-144 (stable sort by score, descending)."""
    scores = [(n, d.score) for n, d in enumerate(detections)]
    scores.sort(key=lambda t: -float(t[1]))  # python's sort is stable like Rust's sort_by
    assert weighted
    return weighted_non_maximum_suppression(scores, detections, min_suppression_threshold, min_score, clusters)


# ----------------------------------------------------------------------------
# transform.rs
# ----------------------------------------------------------------------------
def detection_letterbox_removal(detections, padding):
    """transform.rs:115-142 (scales computed in f64, applied in f32)."""
    left, top, right, bottom = padding
    h_scale = 1.0 - (left + right)
    v_scale = 1.0 - (top + bottom)
    assert h_scale > np.finfo(np.float64).eps and v_scale > np.finfo(np.float64).eps
    out = []
    for d in detections:
        a = d.data.copy()
        a[:, 0] = (a[:, 0] - f32(left)) / f32(h_scale)
        a[:, 1] = (a[:, 1] - f32(top)) / f32(v_scale)
        out.append(Detection(a.astype(np.float32), d.score, d.anchor))
    return out


def select_roi_size(bbox: BBox, image_size, size_mode):
    """transform.rs:87-109."""
    ab = bbox.absolute(image_size)
    width, height = ab.width(), ab.height()
    iw, ih = float(image_size[0]), float(image_size[1])
    if size_mode == SQUARE_LONG:
        long_size = max(width, height)
        return long_size / iw, long_size / ih
    if size_mode == SQUARE_SHORT:
        short = min(width, height)
        return short / iw, short / ih
    return width, height


def bbox_to_roi(bbox: BBox, image_size, rotation_keypoints=None, scale=(1.0, 1.0), size_mode=SIZE_DEFAULT):
    """transform.rs:44-85."""
    if not bbox.normalized():
        raise ValueError("bbox must be normalized")
    width, height = select_roi_size(bbox, image_size, size_mode)
    width, height = width * scale[0], height * scale[1]
    cx = bbox.xmin + bbox.width() / 2.0
    cy = bbox.ymin + bbox.height() / 2.0
    rotation = 0.0
    if rotation_keypoints is not None and len(rotation_keypoints) >= 2:
        (x0, y0), (x1, y1) = rotation_keypoints[0], rotation_keypoints[1]
        angle = -math.atan2(y0 - y1, x1 - x0)
        two_pi = 2.0 * math.pi
        rotation = angle - two_pi * math.floor((angle + math.pi) / two_pi)
    return Rect(cx, cy, width, height, rotation, True)


def bbox_from_landmarks(landmarks):
    """transform.rs:146-165."""
    if len(landmarks) < 2:
        raise ValueError("landmarks must contain at least 2 items")
    xs = [l.x for l in landmarks]
    ys = [l.y for l in landmarks]
    return BBox(min(xs), min(ys), max(xs), max(ys))


@dataclass
class ImageTensor:  # types.rs:5-22
    tensor_data: np.ndarray
    padding: tuple
    original_size: tuple
    u8: np.ndarray = None  # oracle extra: the uint8 image just before normalisation


def image_to_tensor(image: np.ndarray, roi, output_size, keep_aspect_ratio, output_range, flip_horizontal,
                    use_cv2=True) -> ImageTensor:
    """transform.rs:188-309.  ``image``: HxWx3 uint8 RGB.  ``use_cv2=False`` runs the
    numpy restatements of cv_ops instead of cv2 (identical output, slower)."""
    ih, iw = image.shape[:2]
    if roi is None:
        roi = Rect(0.5, 0.5, 1.0, 1.0, 0.0, True)
    roi = roi.scaled((float(iw), float(ih)), False)
    if output_size is None:
        output_size = (int(roi.width), int(roi.height))
    if keep_aspect_ratio:
        width, height = int(roi.size()[0]), int(roi.size()[1])
    else:
        width, height = output_size
    src = np.array(roi.points(), dtype=np.float64).astype(np.float32)
    dst = np.array([(0, 0), (width, 0), (width, height), (0, height)], np.float32)
    if use_cv2:
        m = cv2.getPerspectiveTransform(src, dst, cv2.DECOMP_SVD)  # solveMethod=INTER_LINEAR==1==DECOMP_SVD
        roi_image = cv2.warpPerspective(image, m, (width, height), flags=cv2.INTER_LINEAR,
                                        borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        resize = lambda im, sz: cv2.resize(im, sz, interpolation=cv2.INTER_LINEAR)
    else:
        m = cv_ops.get_perspective_transform_ge(src, dst)
        roi_image = cv_ops.warp_perspective_u8(image, m, (width, height))
        resize = cv_ops.resize_linear_u8
    pad_x = pad_y = 0.0
    if keep_aspect_ratio:
        out_aspect = float(output_size[1] // output_size[0])  # integer division (sic) :240
        roi_aspect = roi.height / roi.width
        new_width, new_height = int(roi.width), int(roi.height)
        if out_aspect > roi_aspect:
            new_height = int(roi.width * out_aspect)
            pad_y = (1.0 - roi_aspect / out_aspect) / 2.0
        else:
            new_width = int(roi.height / out_aspect)
            pad_x = (1.0 - out_aspect / roi_aspect) / 2.0
        if new_width != int(roi.width) or new_height != int(roi.height):
            pad_h, pad_v = int(pad_x * new_width), int(pad_y * new_height)
            padded = cv_ops.copy_make_border_const0(roi_image, pad_v, pad_v, pad_h, pad_h)
            roi_image = resize(padded, (new_width, new_height))
        roi_image = resize(roi_image, (output_size[0], output_size[1]))
    if flip_horizontal:
        roi_image = cv_ops.flip_horizontal(roi_image)
    min_val, max_val = output_range
    # :292-301 (w/h swapped in the reference loops; all outputs are square)
    tensor = (roi_image.astype(np.float64) * (max_val - min_val) / 255.0 + min_val).astype(np.float32)
    return ImageTensor(tensor, (pad_x, pad_y, pad_x, pad_y), (iw, ih), np.ascontiguousarray(roi_image))


def project_landmarks(data: np.ndarray, tensor_size, image_size, padding, roi, flip_horizontal):
    """transform.rs:351-432.  Returns [K,3] float64 (values are f32-rounded)."""
    pts = np.asarray(data, np.float32).reshape(-1, 3).copy()
    width, height = tensor_size
    pts[:, 0] = pts[:, 0] / f32(width)
    pts[:, 1] = pts[:, 1] / f32(height)
    pts[:, 2] = pts[:, 2] / f32(width)
    if flip_horizontal:
        pts[:, 0] = pts[:, 0] * f32(-1.0) + f32(1.0)
    if tuple(padding) != (0.0, 0.0, 0.0, 0.0):
        left, top, right, bottom = padding
        h_scale = 1.0 - (left + right)
        v_scale = 1.0 - (top + bottom)
        pts[:, 0] = ((pts[:, 0].astype(np.float64) - left) / h_scale).astype(np.float32)
        pts[:, 1] = ((pts[:, 1].astype(np.float64) - top) / v_scale).astype(np.float32)
        pts[:, 2] = ((pts[:, 2].astype(np.float64) - 0.0) / h_scale).astype(np.float32)
    if roi is not None:
        nr = roi.scaled((float(image_size[0]), float(image_size[1])), True)
        s, c = f32(math.sin(nr.rotation)), f32(math.cos(nr.rotation))
        x = pts[:, 0] - f32(0.5)
        y = pts[:, 1] - f32(0.5)
        # [x, y, 0] . [[c, s, 0], [-s, c, 0], [1, 1, 1]]   (:393-402)
        rx = (x * c + y * (-s)).astype(np.float32)
        ry = (x * s + y * c).astype(np.float32)
        pts[:, 0] = (rx.astype(np.float64) * nr.width + nr.x_center).astype(np.float32)
        pts[:, 1] = (ry.astype(np.float64) * nr.height + nr.y_center).astype(np.float32)
        pts[:, 2] = (pts[:, 2].astype(np.float64) * nr.width + 0.0).astype(np.float32)
    return pts.astype(np.float64)


# ----------------------------------------------------------------------------
# face_landmark.rs / iris_landmark.rs ROI helpers
# ----------------------------------------------------------------------------
FACE_ROI_SCALE = (1.5, 1.5)       # face_landmark.rs:30
IRIS_ROI_SCALE = (2.3, 2.3)       # iris_landmark.rs:27
LEFT_EYE_START, LEFT_EYE_END, RIGHT_EYE_START, RIGHT_EYE_END = 33, 133, 362, 263  # iris_landmark.rs:29-35
DETECTION_THRESHOLD = f32(0.5)    # face_landmark.rs:31


def face_detection_to_roi(det: Detection, image_size, size_mode=None) -> Rect:
    """face_landmark.rs:180-198."""
    scale = np.array([[f32(image_size[0]), f32(image_size[1])]], np.float32)
    ab = (det.data * scale).astype(np.float32)  # types.rs:237-245
    left_eye = (float(ab[2, 0]), float(ab[2, 1]))
    right_eye = (float(ab[3, 0]), float(ab[3, 1]))
    mode = SQUARE_LONG if size_mode is None else size_mode
    return bbox_to_roi(det.bbox(), image_size, [left_eye, right_eye], FACE_ROI_SCALE, mode)


def iris_roi_from_face_landmarks(lmks: np.ndarray, image_size):
    """iris_landmark.rs:268-292.  lmks [468,3] f64."""
    out = []
    for a, b in ((LEFT_EYE_START, LEFT_EYE_END), (RIGHT_EYE_START, RIGHT_EYE_END)):
        two = [Landmark(*lmks[a]), Landmark(*lmks[b])]
        bbox = bbox_from_landmarks(two)
        kps = [(l.x, l.y) for l in two]
        out.append(bbox_to_roi(bbox, image_size, kps, IRIS_ROI_SCALE, SQUARE_LONG))
    return out[0], out[1]


# --------------------------------------------------------------------------------------------------
# Iris refinement: iris_landmark.rs:64-95 index maps, :380-433 functions.
LEFT_EYE_TO_FACE_LANDMARK_INDEX = np.array([  # iris_landmark.rs:64-78
    33, 7, 163, 144, 145, 153, 154, 155, 133, 246, 161, 160, 159, 158, 157, 173,
    130, 25, 110, 24, 23, 22, 26, 112, 243, 247, 30, 29, 27, 28, 56, 190,
    226, 31, 228, 229, 230, 231, 232, 233, 244, 113, 225, 224, 223, 222, 221, 189,
    35, 124, 46, 53, 52, 65, 143, 111, 117, 118, 119, 120, 121, 128, 245,
    156, 70, 63, 105, 66, 107, 55, 193], np.int32)
RIGHT_EYE_TO_FACE_LANDMARK_INDEX = np.array([  # iris_landmark.rs:80-95
    263, 249, 390, 373, 374, 380, 381, 382, 362, 466, 388, 387, 386, 385, 384, 398,
    359, 255, 339, 254, 253, 252, 256, 341, 463, 467, 260, 259, 257, 258, 286, 414,
    446, 261, 448, 449, 450, 451, 452, 453, 464, 342, 445, 444, 443, 442, 441, 413,
    265, 353, 276, 283, 282, 295, 372, 340, 346, 347, 348, 349, 350, 357, 465,
    383, 300, 293, 334, 296, 336, 285, 417], np.int32)
IRIS_SIZE_IN_MM = 11.8            # iris_landmark.rs:100
IRIS_CENTER, IRIS_LEFT, IRIS_TOP, IRIS_RIGHT, IRIS_BOTTOM = range(5)   # iris_landmark.rs:104-110 `IrisIndex`
NUM_FACE_LANDMARKS = 468


def update_face_landmarks_with_iris_results(face_landmarks, left_contour, right_contour):
    """iris_landmark.rs:380-398: contour point n of each eye replaces face landmark INDEX[n] (x, y and z)."""
    face_landmarks = np.asarray(face_landmarks, np.float64)
    if face_landmarks.shape[0] != NUM_FACE_LANDMARKS:
        raise ValueError("unexpected number of items in face_landmarks")
    refined = face_landmarks.copy()
    left_contour, right_contour = np.asarray(left_contour, np.float64), np.asarray(right_contour, np.float64)
    refined[LEFT_EYE_TO_FACE_LANDMARK_INDEX[:len(left_contour)]] = left_contour
    refined[RIGHT_EYE_TO_FACE_LANDMARK_INDEX[:len(right_contour)]] = right_contour
    return refined


def get_iris_diameter(iris_landmarks, image_size) -> float:
    """iris_landmark.rs:401-418: mean of the horizontal (Left-Right) and vertical (Top-Bottom) extents in pixels, f64."""
    w, h = image_size
    p = np.asarray(iris_landmarks, np.float64)

    def dist(a, b):
        x0, y0, x1, y1 = float(a[0]) * float(w), float(a[1]) * float(h), float(b[0]) * float(w), float(b[1]) * float(h)
        dx, dy = x0 - x1, y0 - y1
        return math.sqrt(dx * dx + dy * dy)        # f64::powi(2) is an exact multiply (numpy's scalar ** 2 is not always)

    return (dist(p[IRIS_TOP], p[IRIS_BOTTOM]) + dist(p[IRIS_LEFT], p[IRIS_RIGHT])) / 2.0


def get_iris_depth(iris_landmarks, focal_length_mm: float, iris_size_px: float, image_size) -> float:
    """iris_landmark.rs:421-433.  The image centre is (width / 2, height / 2) with INTEGER division (:426)."""
    w, h = image_size
    c = np.asarray(iris_landmarks, np.float64)[IRIS_CENTER]
    x0, y0 = float(int(w) // 2), float(int(h) // 2)
    x1, y1 = float(c[0]) * float(w), float(c[1]) * float(h)
    dx, dy, f = x0 - x1, y0 - y1, float(focal_length_mm)
    y = math.sqrt(dx * dx + dy * dy)
    x = math.sqrt(f * f + y * y)
    return IRIS_SIZE_IN_MM * x / float(iris_size_px)
