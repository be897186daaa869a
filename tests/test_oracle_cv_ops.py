"""K4: the integer-exact numpy restatements of the OpenCV ops in oracle/cv_ops.py against cv2,
bit for bit, plus the host build of csrc/glue_math.h (the code the CUDA kernels run) against cv2."""
import ctypes as C
import math

import cv2
import numpy as np
import pytest

from conftest import rng


@pytest.mark.parametrize("src,dst", [((540, 540), 256), ((1920, 1920), 256), ((1920, 1920), 128), ((540, 540), 192), ((57, 57), 64), ((33, 33), 64),
                                     ((63, 63), 64), ((23, 17), 64), ((61, 70), 64), ((128, 128), 64), ((130, 130), 65)])
def test_resize_bit_exact(src, dst):
    from oracle import cv_ops
    img = rng(src[0] * 7 + dst).integers(0, 256, (src[1], src[0], 3), dtype=np.uint8)
    ref = cv2.resize(img, (dst, dst), interpolation=cv2.INTER_LINEAR)
    np.testing.assert_array_equal(cv_ops.resize_linear_u8(img, (dst, dst)), ref)


def test_warp_bit_exact(man):
    from oracle import cv_ops, glue
    r = rng(3)
    for k in range(6):
        roi = glue.Rect(r.uniform(0.2, 0.8), r.uniform(0.2, 0.8), r.uniform(0.2, 0.9), r.uniform(0.2, 0.9), r.uniform(-math.pi, math.pi), True)
        roi = roi.scaled((540.0, 360.0), False)
        src = np.array(roi.points(), np.float64).astype(np.float32)
        size = (192, 192) if k % 2 else (57, 61)
        dst = np.array([(0, 0), (size[0], 0), (size[0], size[1]), (0, size[1])], np.float32)
        m = cv2.getPerspectiveTransform(src, dst, cv2.DECOMP_SVD)
        ref = cv2.warpPerspective(man, m, size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        np.testing.assert_array_equal(cv_ops.warp_perspective_u8(man, m, size), ref)
        mg = cv_ops.get_perspective_transform_ge(src, dst)
        np.testing.assert_allclose(mg, m, rtol=1e-9, atol=1e-9)


def test_image_to_tensor_numpy_equals_cv2(man):
    from oracle import glue
    roi = glue.Rect(0.5, 0.4, 0.4, 0.6, -0.03, True)
    for args in ((None, (256, 256), True, (-1.0, 1.0), False), (roi, (192, 192), False, (0.0, 1.0), False), (roi, (64, 64), True, (0.0, 1.0), True)):
        a = glue.image_to_tensor(man, *args, use_cv2=True)
        b = glue.image_to_tensor(man, *args, use_cv2=False)
        assert (a.u8 != b.u8).mean() <= 1e-4
        assert a.padding == b.padding


# ---- the product's own arithmetic (csrc/glue_math.h compiled for the host) against cv2 ----------------
@pytest.fixture(scope="module")
def hc():
    import hostcheck
    lib = hostcheck.load()
    return lib


def _hc_i2t(hc, image, roi, size, keep, rng_, flip):
    from rs_face_detection_tflite_b200._lib import CRect
    h, w = image.shape[:2]
    out = np.empty((size[1], size[0], 3), np.float32)
    u8 = np.empty((size[1], size[0], 3), np.uint8)
    pad = (C.c_double * 4)()
    croi = None
    if roi is not None:
        croi = CRect(roi.x_center, roi.y_center, roi.width, roi.height, roi.rotation, 1 if roi.normalized else 0, 0)
    img = np.ascontiguousarray(image)
    rc = hc.hc_image_to_tensor(img.ctypes.data_as(C.c_void_p), w, h, C.byref(croi) if croi is not None else None, size[0], size[1], int(keep),
                               C.c_double(rng_[0]), C.c_double(rng_[1]), int(flip), out.ctypes.data_as(C.c_void_p), u8.ctypes.data_as(C.c_void_p), pad)
    assert rc == 0
    return out, u8, tuple(pad)


def test_kernel_math_letterbox_bit_exact(hc, man):
    from oracle import glue
    import synth_frames
    for img, size in ((man, 256), (man, 128), (synth_frames.noise_frames(1, 640, 480, 3)[0], 192), (synth_frames.face_frame(0), 256)):
        ref = glue.image_to_tensor(img, None, (size, size), True, (-1.0, 1.0), False)
        t, u8, pad = _hc_i2t(hc, img, None, (size, size), True, (-1.0, 1.0), False)
        np.testing.assert_array_equal(u8, ref.u8)
        np.testing.assert_array_equal(t, ref.tensor_data)
        assert pad == tuple(ref.padding)


def test_kernel_math_warps_match_opencv(hc, man):
    from oracle import glue
    r = rng(8)
    for k in range(8):
        roi = glue.Rect(r.uniform(0.3, 0.7), r.uniform(0.3, 0.7), r.uniform(0.1, 0.8), r.uniform(0.1, 0.8), r.uniform(-3, 3), True)
        if k % 2:
            side = r.uniform(20, 200)
            roi = glue.Rect(roi.x_center, roi.y_center, side / 540, side / 360, roi.rotation, True)
            args = ((64, 64), True, (0.0, 1.0), bool(k & 2))
        else:
            args = ((192, 192), False, (0.0, 1.0), False)
        ref = glue.image_to_tensor(man, roi, *args)
        t, u8, pad = _hc_i2t(hc, man, roi, *args)
        d = np.abs(u8.astype(int) - ref.u8.astype(int))
        assert d.max() <= 1 and (d > 0).mean() <= 2e-3
        assert pad == tuple(ref.padding)
