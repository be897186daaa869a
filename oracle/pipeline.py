"""Oracle mirror of the reference's public API (CPU; cv2 + torch-CPU + numpy).

Test infrastructure; see ``oracle/__init__.py``.  Mirrors
``FaceDetection::new/infer`` (face_detection.rs:153-267), ``FaceLandmark::new/
infer`` (face_landmark.rs:208-306), ``IrisLandmark::new/infer``
(iris_landmark.rs:142-248) and the canonical call sequence of lib.rs:20-40.
"""
from __future__ import annotations

import os

import numpy as np

from . import glue
from .graph_exec import GraphExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_MODEL_DIR = os.path.join(os.path.dirname(_HERE), "models")


class FaceDetection:
    def __init__(self, model_type=glue.BACK_CAMERA, model_dir=None):
        if model_type not in glue.MODEL_FILES:
            raise ValueError("unsupported model type")
        self.model_type = model_type
        self.net = GraphExecutor(os.path.join(model_dir or DEFAULT_MODEL_DIR, glue.MODEL_FILES[model_type]))
        self.anchors = glue.ssd_generate_anchors(model_type)
        self.height, self.width = self.net.input_shape[1], self.net.input_shape[2]

    def preprocess(self, image, roi=None):
        return glue.image_to_tensor(image, roi, (self.width, self.height), True, (-1.0, 1.0), False)

    def forward(self, tensor):
        """tensor [B,S,S,3] -> (regressors [B,N,16], classificators [B,N,1])."""
        reg, cls = self.net.run(tensor)
        return reg, cls

    def postprocess(self, reg, cls, padding, trace=None):
        """One frame: raw [N,16],[N,1] -> list[Detection] (face_detection.rs:259-266)."""
        boxes = glue.decode_boxes(reg, self.anchors, float(self.height))
        scores = glue.get_sigmoid_score(cls)
        dets = glue.convert_to_detections(boxes, scores)
        clusters = []
        pruned = glue.non_maximum_suppression(dets, clusters=clusters)
        if trace is not None:
            trace["survivors"] = [d.anchor for d in dets]
            trace["clusters"] = clusters
            trace["scores"] = scores.reshape(-1)
        return glue.detection_letterbox_removal(pruned, padding)

    def infer(self, image, roi=None, trace=None):
        it = self.preprocess(image, roi)
        reg, cls = self.forward(it.tensor_data[None])
        if trace is not None:
            trace["tensor"] = it.tensor_data
            trace["u8"] = it.u8
            trace["padding"] = it.padding
            trace["regressors"] = reg[0]
            trace["classificators"] = cls[0]
        return self.postprocess(reg[0], cls[0], it.padding, trace)


class FaceLandmark:
    def __init__(self, model_path=None):
        self.net = GraphExecutor(model_path or os.path.join(DEFAULT_MODEL_DIR, "face_landmark.tflite"))
        self.height, self.width = self.net.input_shape[1], self.net.input_shape[2]

    def preprocess(self, image, roi):
        return glue.image_to_tensor(image, roi, (self.width, self.height), False, (0.0, 1.0), False)

    def infer(self, image, roi=None, trace=None):
        it = self.preprocess(image, roi)
        raw, flag = self.net.run(it.tensor_data[None])
        if trace is not None:
            trace.update(tensor=it.tensor_data, u8=it.u8, raw=raw.reshape(-1), flag=float(flag.reshape(-1)[-1]))
        face_flag = glue.sigmoid_f32(flag.reshape(-1))[-1]
        if face_flag <= glue.DETECTION_THRESHOLD:
            return np.zeros((0, 3), np.float64)
        return glue.project_landmarks(raw, (self.width, self.height), it.original_size, it.padding, roi, False)


class IrisLandmark:
    def __init__(self, model_path=None):
        self.net = GraphExecutor(model_path or os.path.join(DEFAULT_MODEL_DIR, "iris_landmark.tflite"))
        self.height, self.width = self.net.input_shape[1], self.net.input_shape[2]

    def preprocess(self, image, roi, is_right_eye):
        return glue.image_to_tensor(image, roi, (self.width, self.height), True, (0.0, 1.0), is_right_eye)

    def infer(self, image, roi=None, is_right_eye=False, trace=None):
        it = self.preprocess(image, roi, is_right_eye)
        eye, iris = self.net.run(it.tensor_data[None])
        if trace is not None:
            trace.update(tensor=it.tensor_data, u8=it.u8, raw_eye=eye.reshape(-1), raw_iris=iris.reshape(-1),
                         padding=it.padding)
        contour = glue.project_landmarks(eye, (self.width, self.height), it.original_size, it.padding, roi,
                                         is_right_eye)
        irisl = glue.project_landmarks(iris, (self.width, self.height), it.original_size, it.padding, roi,
                                       is_right_eye)
        return contour, irisl


class Pipeline:
    """detect -> face ROI -> landmark -> eye ROIs -> iris(L,R), as lib.rs:20-40."""

    def __init__(self, model_type=glue.BACK_CAMERA, model_dir=None):
        d = model_dir or DEFAULT_MODEL_DIR
        self.det = FaceDetection(model_type, d)
        self.lmk = FaceLandmark(os.path.join(d, "face_landmark.tflite"))
        self.iris = IrisLandmark(os.path.join(d, "iris_landmark.tflite"))

    def run(self, image, max_faces=1):
        h, w = image.shape[:2]
        faces = self.det.infer(image)
        out = []
        for face in faces[:max_faces]:
            roi = glue.face_detection_to_roi(face, (w, h))
            lm = self.lmk.infer(image, roi)
            rec = dict(detection=face, roi=roi, landmarks=lm)
            if len(lm):
                lroi, rroi = glue.iris_roi_from_face_landmarks(lm, (w, h))
                rec["left_roi"], rec["right_roi"] = lroi, rroi
                rec["right"] = self.iris.infer(image, rroi, True)
                rec["left"] = self.iris.infer(image, lroi, False)
            out.append(rec)
        return faces, out
