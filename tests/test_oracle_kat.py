"""The oracle against every golden artefact the reference holds for this path:
K1 the three rendered PNGs in assets/ written by src/lib.rs:42-83 (pixel-exact), K2 the survey's recorded
values on test_data/man.jpg, K3 the anchor-table hashes.  CPU only."""
import hashlib
import os

import cv2
import numpy as np
import pytest

from conftest import MODELS, ROOT


@pytest.fixture(scope="module")
def run(man, oracle_pipeline):
    from oracle import glue
    p = oracle_pipeline[glue.BACK_CAMERA]
    tr = {}
    faces = p.det.infer(man, trace=tr)
    faces2, out = p.run(man)
    return dict(faces=faces, trace=tr, out=out[0])


def _asset(name):
    bgr = cv2.imread(os.path.join(ROOT, "assets", name), cv2.IMREAD_COLOR)
    return bgr[:, :, ::-1].astype(int)


def test_input_image_hash(man):
    assert man.shape == (360, 540, 3)
    assert hashlib.sha256(man.tobytes()).hexdigest()[:16] == "5c193162f846f56e"


def test_k3_anchor_tables():
    from oracle import glue
    for model, sha, n in ((0, "7527f7bf39f988ed", 896), (1, "7527f7bf39f988ed", 896), (2, "7527f7bf39f988ed", 896), (3, "30f2249dab7a1447", 2304)):
        a = glue.ssd_generate_anchors(model)
        assert a.shape == (n, 2) and a.dtype == np.float32
        assert hashlib.sha256(a.tobytes()).hexdigest()[:16] == sha


def test_k1_bbox_png_pixel_exact(run):
    """assets/man_bbox.png: pure-green hollow rectangle drawn by render.rs:446-461 at
    Rect::at(int(xmin*W), int(ymin*H)).of_size(int(w*W), int(h*H))."""
    img = _asset("man_bbox.png")
    green = (img[:, :, 0] == 0) & (img[:, :, 1] == 255) & (img[:, :, 2] == 0)
    ys, xs = np.nonzero(green)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (195, 333, 74, 212)
    b = run["faces"][0].bbox()
    x0, y0, w, h = int(b.xmin * 540), int(b.ymin * 360), int(b.width() * 540), int(b.height() * 360)
    assert (x0, y0, w, h) == (195, 74, 139, 139)
    assert (x0 + w - 1, y0 + h - 1) == (333, 212)


def test_k1_landmark_png_span(run):
    """assets/man_landmark.png: 468 red 2x2 points at (int(x)-1, int(y)-1) (render.rs:424-427) plus connection lines."""
    img = _asset("man_landmark.png")
    red = (img[:, :, 0] == 255) & (img[:, :, 1] == 0) & (img[:, :, 2] == 0)
    ys, xs = np.nonzero(red)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (201, 324, 65, 209)
    lm = run["out"]["landmarks"]
    px = lm[:, 0] * 540
    py = lm[:, 1] * 360
    assert int(px.min()) - 1 == 201 and int(py.min()) - 1 == 65
    # every rendered landmark pixel is red in the asset
    hits = sum(bool(red[int(y) - 1, int(x) - 1]) for x, y in zip(px, py))
    assert hits == 468


def test_k1_iris_png_points(run):
    """assets/man_iris.png: the 15 eyeball-contour points of both eyes."""
    img = _asset("man_iris.png")
    red = (img[:, :, 0] == 255) & (img[:, :, 1] == 0) & (img[:, :, 2] == 0)
    ys, xs = np.nonzero(red)
    assert (xs.min(), xs.max(), ys.min(), ys.max()) == (224, 301, 104, 111)
    hits = total = 0
    for key in ("left", "right"):
        contour = run["out"][key][0][:15]
        for x, y in zip(contour[:, 0] * 540, contour[:, 1] * 360):
            total += 1
            hits += bool(red[int(y) - 1, int(x) - 1])
    assert hits == total == 30


def test_k2_recorded_values(run):
    tr = run["trace"]
    assert tr["survivors"] == [207, 209, 239, 241]
    assert tr["clusters"] == [[239, 209, 207, 241]]
    logits = tr["classificators"].reshape(-1)[[207, 209, 239, 241]]
    np.testing.assert_allclose(logits, [2.673951, 2.818204, 3.165867, 2.232641], atol=2e-4)
    f = run["faces"][0]
    assert abs(float(f.score) - 0.959529) < 1e-5
    np.testing.assert_allclose(f.data, [[0.362137, 0.205837], [0.620166, 0.592880], [0.439248, 0.313293], [0.541184, 0.309208],
                                        [0.490653, 0.406025], [0.491580, 0.481972], [0.382182, 0.353412], [0.599753, 0.346908]], atol=2e-6)
    assert tr["padding"] == (0.0, pytest.approx(1 / 6, abs=1e-12), 0.0, pytest.approx(1 / 6, abs=1e-12))
    t = tr["tensor"]
    assert (t[:42] == -1.0).all() and (t[214:] == -1.0).all() and not (t[42] == -1.0).all()
    roi = run["out"]["roi"]
    np.testing.assert_allclose([roi.x_center, roi.y_center, roi.width, roi.height, roi.rotation],
                               [0.491151, 0.399359, 0.387043, 0.580565, -0.0267066], atol=2e-6)
    lroi, rroi = run["out"]["left_roi"], run["out"]["right_roi"]
    np.testing.assert_allclose([lroi.x_center, lroi.y_center, lroi.width, lroi.height, lroi.rotation], [0.43417, 0.30605, 0.10598, 0.15898, 0.02574], atol=2e-5)
    np.testing.assert_allclose([rroi.x_center, rroi.y_center, rroi.width, rroi.height, rroi.rotation], [0.53808, 0.30213, 0.10515, 0.15773, -0.10522], atol=2e-5)


@pytest.mark.parametrize("model,surv,score", [(2, [206, 207, 208, 209, 238, 239, 240, 241], 0.932164), (3, [983], 0.934486)])
def test_k2_other_detectors(man, model, surv, score):
    from oracle import pipeline
    tr = {}
    faces = pipeline.FaceDetection(model, MODELS).infer(man, trace=tr)
    assert tr["survivors"] == surv
    assert abs(float(faces[0].score) - score) < 2e-5
