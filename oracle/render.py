"""oracle/render.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): restatement of the drawing the reference used to write its
only result artefacts, assets/man_{bbox,landmark,iris}.png (src/lib.rs:42-83):

    detections_to_render_data / landmarks_to_render_data   render.rs:262-359
    render_to_image                                         render.rs:361-479
    FACE_LANDMARK_CONNECTIONS, EYE_LANDMARK_CONNECTIONS     face_landmark.rs:35-160, iris_landmark.rs:44-60 (data tables)

With it the oracle is pinned not to a few facts about those PNGs but to EVERY drawn pixel: tests/test_golden.py renders the
oracle's detection / 468 landmarks / eye contours on test_data/man.jpg and requires the drawn pixel set to equal the reference's.
The pixel routines live in the un-vendored `imageproc` crate (0.25.0, Cargo.lock:547-549); restated from its published source:
`draw_filled_rect_mut` / `draw_hollow_rect_mut` clip to the image, `draw_line_segment_mut` walks `BresenhamLineIter` (steep lines
transposed, start/end ordered by x, error term dx/2 decremented by |dy|, y stepped when it goes negative) and skips pixels
outside the image.  SURVEY.md 8f rank 4; nothing in the product path draws.
"""
from __future__ import annotations

import numpy as np

FACE_LANDMARK_CONNECTIONS = [
    (61, 146), (146, 91), (91, 181), (181, 84), (84, 17), (17, 314), (314, 405), (405, 321), (321, 375), (375, 291),
    (61, 185), (185, 40), (40, 39), (39, 37), (37, 0), (0, 267), (267, 269), (269, 270), (270, 409), (409, 291),
    (78, 95), (95, 88), (88, 178), (178, 87), (87, 14), (14, 317), (317, 402), (402, 318), (318, 324), (324, 308),
    (78, 191), (191, 80), (80, 81), (81, 82), (82, 13), (13, 312), (312, 311), (311, 310), (310, 415), (415, 308),
    (33, 7), (7, 163), (163, 144), (144, 145), (145, 153), (153, 154), (154, 155), (155, 133), (33, 246), (246, 161),
    (161, 160), (160, 159), (159, 158), (158, 157), (157, 173), (173, 133), (46, 53), (53, 52), (52, 65), (65, 55),
    (70, 63), (63, 105), (105, 66), (66, 107), (263, 249), (249, 390), (390, 373), (373, 374), (374, 380), (380, 381),
    (381, 382), (382, 362), (263, 466), (466, 388), (388, 387), (387, 386), (386, 385), (385, 384), (384, 398), (398, 362),
    (276, 283), (283, 282), (282, 295), (295, 285), (300, 293), (293, 334), (334, 296), (296, 336), (10, 338), (338, 297),
    (297, 332), (332, 284), (284, 251), (251, 389), (389, 356), (356, 454), (454, 323), (323, 361), (361, 288), (288, 397),
    (397, 365), (365, 379), (379, 378), (378, 400), (400, 377), (377, 152), (152, 148), (148, 176), (176, 149), (149, 150),
    (150, 136), (136, 172), (172, 58), (58, 132), (132, 93), (93, 234), (234, 127), (127, 162), (162, 21), (21, 54),
    (54, 103), (103, 67), (67, 109), (109, 10),
]
EYE_LANDMARK_CONNECTIONS = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (9, 10), (10, 11), (11, 12), (12, 13), (13, 14),
                            (0, 9), (8, 14)]
MAX_EYE_LANDMARK = len(EYE_LANDMARK_CONNECTIONS)


def _bresenham(mask: np.ndarray, x0: float, y0: float, x1: float, y1: float) -> None:
    """imageproc::drawing::draw_line_segment_mut on a boolean canvas."""
    h, w = mask.shape
    steep = abs(y1 - y0) > abs(x1 - x0)
    if steep:
        x0, y0, x1, y1 = y0, x0, y1, x1
    if x0 > x1:
        x0, x1, y0, y1 = x1, x0, y1, y0
    dx = np.float32(x1 - x0)
    dy = np.float32(abs(y1 - y0))
    err = np.float32(dx / np.float32(2))
    x, y, end_x = int(x0), int(y0), int(x1)
    y_step = 1 if y0 < y1 else -1
    while x <= end_x:
        px, py = (y, x) if steep else (x, y)
        if 0 <= px < w and 0 <= py < h:
            mask[py, px] = True
        x += 1
        err = np.float32(err - dy)
        if err < 0:
            y += y_step
            err = np.float32(err + dx)


def _filled_rect(mask: np.ndarray, left: int, top: int, rw: int, rh: int) -> None:
    h, w = mask.shape
    mask[max(top, 0):max(min(top + rh, h), 0), max(left, 0):max(min(left + rw, w), 0)] = True


def _hollow_rect(mask: np.ndarray, left: int, top: int, rw: int, rh: int) -> None:
    """draw_hollow_rect_mut: the four edges of Rect::at(left, top).of_size(rw, rh) as line segments (right = left + rw - 1)."""
    right, bottom = left + rw - 1, top + rh - 1
    _bresenham(mask, left, top, right, top)
    _bresenham(mask, left, bottom, right, bottom)
    _bresenham(mask, left, top, left, bottom)
    _bresenham(mask, right, top, right, bottom)


def render_landmarks_mask(landmarks, connections, size, thickness: float = 2.0) -> np.ndarray:
    """Pixels painted by landmarks_to_render_data(normalized) + render_to_image: the connection lines, then the points
    (render.rs:315-359, :404-427).  landmarks: [n, >=2] normalised; size = (width, height)."""
    w, h = size
    lm = np.asarray(landmarks, np.float64)
    mask = np.zeros((h, w), bool)
    for a, b in connections:
        xs, ys, xe, ye = lm[a, 0] * w, lm[a, 1] * h, lm[b, 0] * w, lm[b, 1] * h
        _bresenham(mask, float(int(xs)), float(int(ys)), float(int(xe)), float(int(ye)))      # `as i32` then `as f32`
    t = int(thickness)
    pw = max(t // 2, 1)
    for x, y in zip(lm[:, 0] * w, lm[:, 1] * h):
        _filled_rect(mask, int(x) - pw, int(y) - pw, 2 * pw, 2 * pw)                           # `as u32` of a non-negative value
    return mask


def render_detection_mask(bbox_xyxy, size) -> np.ndarray:
    """Pixels painted for one detection's bounds by detections_to_render_data(bounds only) + render_to_image (render.rs:446-461).
    (The `line_width` the caller passes ends up unused by draw_hollow_rect_mut: the outline is one pixel wide.)"""
    w, h = size
    xmin, ymin, xmax, ymax = (float(v) for v in bbox_xyxy)
    left, top, right, bottom = xmin * w, ymin * h, xmax * w, ymax * h
    mask = np.zeros((h, w), bool)
    _hollow_rect(mask, int(left), int(top), int(right - left), int(bottom - top))
    return mask
