"""Synthetic 1080p frame generators shared by tests/ and bench.py.

G1 ``noise_frames``: uniform uint8 noise -- preprocessing / throughput only (no faces).
G2 ``face_frames``:  test_data/*.jpg pasted at a seeded random scale, rotation and position onto a
    seeded smooth background, so that every frame carries exactly one face and the whole
    detect -> landmark -> iris chain has work to do.

Data generation only (cv2 is used to decode the JPEG and to paste it); nothing here is on the product path.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FACES = ("man.jpg", "russ_cox_1.jpg", "russ_cox_2.jpg")
_cache = {}


def load_rgb(name: str) -> np.ndarray:
    """Decode test_data/<name> to RGB uint8 (what utils.rs:8-21 convert_image_to_mat yields)."""
    import cv2
    if name not in _cache:
        bgr = cv2.imread(os.path.join(_HERE, "test_data", name), cv2.IMREAD_COLOR)
        if bgr is None:
            raise FileNotFoundError(name)
        _cache[name] = np.ascontiguousarray(bgr[:, :, ::-1])
    return _cache[name]


def noise_frames(n: int, width: int = 1920, height: int = 1080, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, height, width, 3), dtype=np.uint8)


def face_frame(index: int, width: int = 1920, height: int = 1080, face: str = "man.jpg") -> np.ndarray:
    """One G2 frame; the seed is the frame index."""
    import cv2
    rng = np.random.default_rng(1000 + index)
    src = load_rgb(face)
    sh, sw = src.shape[:2]
    # smooth background: per-channel linear gradient + low-amplitude noise
    yy = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None, None]
    xx = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :, None]
    c0 = rng.uniform(40, 200, 3).astype(np.float32)
    gx = rng.uniform(-40, 40, 3).astype(np.float32)
    gy = rng.uniform(-40, 40, 3).astype(np.float32)
    bg = c0 + gx * xx + gy * yy + rng.normal(0.0, 2.0, (height, width, 3)).astype(np.float32)
    frame = np.clip(bg, 0, 255).astype(np.uint8)
    scale = float(rng.uniform(1.0, 2.5))
    theta = float(rng.uniform(-30.0, 30.0))
    # paste centre so that the rotated, scaled image stays inside the frame
    half = 0.5 * scale * float(np.hypot(sw, sh))
    cx = float(rng.uniform(min(half, width / 2), max(width - half, width / 2)))
    cy = float(rng.uniform(min(half, height / 2), max(height - half, height / 2)))
    m = cv2.getRotationMatrix2D((sw / 2.0, sh / 2.0), theta, scale)
    m[0, 2] += cx - sw / 2.0
    m[1, 2] += cy - sh / 2.0
    warped = cv2.warpAffine(src, m, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    mask = cv2.warpAffine(np.full((sh, sw), 255, np.uint8), m, (width, height), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT,
                          borderValue=0)
    frame[mask > 0] = warped[mask > 0]
    return frame


def face_frames(n: int, width: int = 1920, height: int = 1080, start: int = 0, faces=("man.jpg",)) -> np.ndarray:
    out = np.empty((n, height, width, 3), np.uint8)
    for i in range(n):
        out[i] = face_frame(start + i, width, height, faces[(start + i) % len(faces)])
    return out


def multi_face_frame(index: int, width: int = 1920, height: int = 1080, faces=("man.jpg", "russ_cox_1.jpg")) -> np.ndarray:
    """A G2-style frame that carries len(faces) faces: the frame is cut into equal vertical strips and one test image is pasted
    (seeded scale / rotation / position, seed = frame index) inside each strip, so the faces never overlap.  Exercises the
    max_faces fan-out and multi-cluster NMS of the pipeline."""
    import cv2
    rng = np.random.default_rng(5000 + index)
    n = len(faces)
    yy = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None, None]
    xx = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :, None]
    bg = rng.uniform(60, 180, 3).astype(np.float32) + rng.uniform(-30, 30, 3).astype(np.float32) * xx + rng.uniform(-30, 30, 3).astype(np.float32) * yy
    frame = np.clip(bg + rng.normal(0.0, 2.0, (height, width, 3)).astype(np.float32), 0, 255).astype(np.uint8)
    strip = width // n
    for k, name in enumerate(faces):
        src = load_rgb(name)
        sh, sw = src.shape[:2]
        theta = float(rng.uniform(-20.0, 20.0))
        # largest scale whose rotated bounding circle fits the strip and the frame height
        smax = min(strip, height) / float(np.hypot(sw, sh))
        scale = float(rng.uniform(0.75 * smax, 0.98 * smax))
        half = 0.5 * scale * float(np.hypot(sw, sh))
        cx = k * strip + float(rng.uniform(half, max(strip - half, half)))
        cy = float(rng.uniform(half, max(height - half, half)))
        m = cv2.getRotationMatrix2D((sw / 2.0, sh / 2.0), theta, scale)
        m[0, 2] += cx - sw / 2.0
        m[1, 2] += cy - sh / 2.0
        warped = cv2.warpAffine(src, m, (width, height), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        mask = cv2.warpAffine(np.full((sh, sw), 255, np.uint8), m, (width, height), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT,
                              borderValue=0)
        frame[mask > 0] = warped[mask > 0]
    return frame


def grid_frame(grid: int = 6, cell: int = 256) -> np.ndarray:
    """A square frame tiled with grid x grid copies of the face of man.jpg: more detections than FDL_MAX_DETECTIONS (32) for
    grid >= 6 -- the pipeline's overflow report is tested with it."""
    import cv2
    man = load_rgb("man.jpg")
    crop = cv2.resize(np.ascontiguousarray(man[20:260, 150:390]), (cell - 16, cell - 16))
    frame = np.full((grid * cell, grid * cell, 3), 120, np.uint8)
    for i in range(grid):
        for j in range(grid):
            frame[i * cell + 8:i * cell + cell - 8, j * cell + 8:j * cell + cell - 8] = crop
    return frame
