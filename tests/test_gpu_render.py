"""render.rs on the device: fdl_render_to_image + the annotation builders of the Python mirror against the
reference's own rendered assets.  tests/golden/ holds the exact pixel sets of assets/man_{bbox,landmark,iris}.png and the oracle's
vectors on man.jpg that (test_golden.py) paint exactly those sets through the oracle's restatement of the drawing code; painted
through the product's kernels they must give the same sets -- 552 + 2414 + 150 pixels, none missing, none extra."""
import os

import numpy as np
import pytest

from conftest import MODELS, ROOT
from test_golden import _mask, gold, k1   # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def _painted(rgba, base, color):
    m = (rgba[:, :, 0] == color[0]) & (rgba[:, :, 1] == color[1]) & (rgba[:, :, 2] == color[2])
    was = (base[:, :, 0] == color[0]) & (base[:, :, 1] == color[1]) & (base[:, :, 2] == color[2])
    return m & ~was


def test_the_three_reference_renders_pixel_for_pixel(fdl, gpu, man, k1, gold):
    import cv2
    lm = [fdl.Landmark(*map(float, p)) for p in gold["landmarks"]]
    det = fdl.Detection(np.asarray(gold["det_data"][0], np.float32).reshape(8, 2), float(gold["det_score"][0]))
    base = man.copy()
    base[(base[:, :, 0] == 255) & (base[:, :, 1] == 0) & (base[:, :, 2] == 0)] = (254, 0, 0)      # no pure red / green in the photo itself
    base[(base[:, :, 0] == 0) & (base[:, :, 1] == 255) & (base[:, :, 2] == 0)] = (0, 254, 0)
    # lib.rs:44-57: the face bounding box
    rgba = fdl.render_to_image(fdl.detections_to_render_data([det], fdl.Colors.GREEN, None, 4, 2, True, None), base, device=gpu)
    assert rgba.shape == (360, 540, 4) and (rgba[:, :, 3] == 255).all()
    img = cv2.imread(os.path.join(ROOT, "assets", "man_bbox.png"), cv2.IMREAD_COLOR)[:, :, ::-1]
    green = (img[:, :, 0] == 0) & (img[:, :, 1] == 255) & (img[:, :, 2] == 0)
    drawn = _painted(rgba, base, (0, 255, 0))
    assert int(drawn.sum()) == k1["bbox_green_pixels"]
    np.testing.assert_array_equal(drawn, green)
    np.testing.assert_array_equal(rgba[~drawn][:, :3], base[~drawn])                              # everything else is the photo
    # lib.rs:59-62: the 468 landmarks and their connections
    rgba = fdl.render_to_image(fdl.face_landmarks_to_render_data(lm, fdl.Colors.RED, fdl.Colors.RED, 2.0, None), base, device=gpu)
    drawn = _painted(rgba, base, (255, 0, 0))
    assert int(drawn.sum()) == k1["landmark_red_pixels"] == 2414
    np.testing.assert_array_equal(drawn, _mask("landmark", k1))
    # lib.rs:65-82: both eyeball contours in one annotation list, right eye first
    ann = []
    for key in ("right", "left"):
        contour = [fdl.Landmark(*map(float, p)) for p in gold[key + "_contour"]]
        ann = fdl.eye_landmarks_to_render_data(fdl.IrisResults(contour, contour[:5]).eyeball_contour(), fdl.Colors.RED, fdl.Colors.RED, 2.0, ann)
    rgba = fdl.render_to_image(ann, base, device=gpu)
    np.testing.assert_array_equal(_painted(rgba, base, (255, 0, 0)), _mask("iris", k1))


def test_paint_order_clipping_and_errors(fdl, gpu):
    """Later annotations paint over earlier ones (no blending); primitives are clipped to the image; an empty rectangle is the
    reference's panic (imageproc Rect::of_size), an absurd coordinate is rejected instead of walked."""
    from oracle import render as orender
    img = np.zeros((40, 60, 3), np.uint8)
    ann = [fdl.Annotation([("filled_rect", 5.0, 5.0, 30.0, 25.0)], False, 1.0, fdl.Colors.BLUE),
           fdl.Annotation([("line", -20.0, 10.0, 80.0, 35.0), ("point", 0.0, 0.0), ("point", 59.9, 39.9)], False, 6.0, fdl.Colors.RED),
           fdl.Annotation([("rect", 0.1, 0.2, 0.9, 0.8)], True, 1.0, fdl.Colors.GREEN)]
    rgba = fdl.render_to_image(ann, img, device=gpu)
    want = np.zeros((40, 60, 3), np.uint8)
    want[5:25, 5:30] = (0, 0, 255)
    m = np.zeros((40, 60), bool)
    orender._bresenham(m, -20.0, 10.0, 80.0, 35.0)
    orender._filled_rect(m, 0 - 3, 0 - 3, 6, 6)
    orender._filled_rect(m, 59 - 3, 39 - 3, 6, 6)
    want[m] = (255, 0, 0)
    m = np.zeros((40, 60), bool)
    orender._hollow_rect(m, int(0.1 * 60), int(0.2 * 40), int(0.9 * 60 - 0.1 * 60), int(0.8 * 40 - 0.2 * 40))
    want[m] = (0, 255, 0)
    np.testing.assert_array_equal(rgba[:, :, :3], want)
    with pytest.raises(fdl.FdlError):
        fdl.render_to_image([fdl.Annotation([("rect", 5.0, 5.0, 5.5, 9.0)], False, 1.0, fdl.Colors.RED)], img, device=gpu)
    with pytest.raises(fdl.FdlError):
        fdl.render_to_image([fdl.Annotation([("line", 0.0, 0.0, 1e12, 5.0)], False, 1.0, fdl.Colors.RED)], img, device=gpu)
    assert fdl.render_to_image([], img, device=gpu)[:, :, :3].sum() == 0


def test_render_of_the_gpu_results_lands_on_the_reference_pixels(fdl, gpu, man, k1):
    """The whole lib.rs flow on the device -- infer, then draw: the rectangle exactly, the landmark drawing within the few pixels a
    < 0.5 px coordinate difference can move."""
    det = fdl.FaceDetection(fdl.FaceDetectionModel.BackCamera, MODELS, device=gpu)
    lmk = fdl.FaceLandmark(MODELS + "/face_landmark.tflite", device=gpu)
    faces = det.infer(man)
    lm = lmk.infer(man, fdl.face_detection_to_roi(faces[0], (540, 360)))
    base = man.copy()
    base[(base[:, :, 0] == 255) & (base[:, :, 1] == 0) & (base[:, :, 2] == 0)] = (254, 0, 0)
    base[(base[:, :, 0] == 0) & (base[:, :, 1] == 255) & (base[:, :, 2] == 0)] = (0, 254, 0)
    rgba = fdl.render_to_image(fdl.detections_to_render_data(faces, fdl.Colors.GREEN, None, 4, 2, True), base, device=gpu)
    x0, x1, y0, y1 = k1["bbox_green_extent_x0x1y0y1"]
    ys, xs = np.nonzero(_painted(rgba, base, (0, 255, 0)))
    assert [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())] == [x0, x1, y0, y1]
    rgba = fdl.render_to_image(fdl.face_landmarks_to_render_data(lm, fdl.Colors.RED, fdl.Colors.RED, 2.0), base, device=gpu)
    drawn, ref = _painted(rgba, base, (255, 0, 0)), _mask("landmark", k1)
    assert (drawn ^ ref).sum() <= 0.03 * ref.sum(), int((drawn ^ ref).sum())
    det.close(); lmk.close()
