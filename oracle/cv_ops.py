"""Integer-exact numpy restatements of the OpenCV ops on the hot path.

Test infrastructure; see ``oracle/__init__.py``.

The reference calls, through crate ``opencv 0.93.1`` (un-vendored; links the system
libopencv): ``get_perspective_transform`` (transform.rs:222), ``warp_perspective``
(:226-234), ``copy_make_border`` (:260-269), ``resize`` (:272, :277) and ``flip``
(:284), all on 8UC3 images with INTER_LINEAR / BORDER_CONSTANT(0).

cv2 4.13 (same library family) is importable here and is the direct oracle for
these ops; the functions below restate OpenCV's *fixed-point* arithmetic
 so that (1) the algorithm the CUDA kernels implement is
written down, and (2) tests can run without trusting a black box.  They are
checked bit-for-bit against cv2 in ``tests/test_oracle_cv_ops.py``.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS           # 32: warp coordinates live on a 1/32 px grid
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS
INTER_RESIZE_COEF_BITS = 11
INTER_RESIZE_COEF_SCALE = 1 << INTER_RESIZE_COEF_BITS


# ----------------------------------------------------------------------------
# resize(..., INTER_LINEAR) on 8U  (transform.rs:272,277)
# ----------------------------------------------------------------------------
def _resize_axis_coeffs(dn: int, sn: int, clamp_frac: bool):
    """Per-destination source index and 11-bit coefficient pair for one axis.

    x axis (clamp_frac=True): when the left tap falls outside, the fraction is
    zeroed and the index clamped.  y axis: the fraction is kept and only the row
    *indices* are clamped into [0, sn-1] when the rows are fetched.
    """
    scale = np.float64(sn) / np.float64(dn)
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    fr = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        fr[lo] = 0.0
        s[lo] = 0
        hi = s >= sn - 1
        fr[hi] = 0.0
        s[hi] = sn - 1
    c0 = np.rint((np.float32(1.0) - fr) * np.float32(INTER_RESIZE_COEF_SCALE)).astype(np.int32)
    c1 = np.rint(fr * np.float32(INTER_RESIZE_COEF_SCALE)).astype(np.int32)
    return s, c0, c1


def resize_linear_u8(src: np.ndarray, dsize) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR) for uint8 HxWxC."""
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape[:2]
    if (dw, dh) == (sw, sh):
        return src.copy()
    xs, a0, a1 = _resize_axis_coeffs(dw, sw, True)
    ys, b0, b1 = _resize_axis_coeffs(dh, sh, False)
    x0 = xs
    x1 = np.minimum(xs + 1, sw - 1)
    y0 = np.clip(ys, 0, sh - 1)
    y1 = np.clip(ys + 1, 0, sh - 1)
    s = src.astype(np.int32)
    # horizontal pass: H[y][dx] = S[y][x0]*a0 + S[y][x1]*a1  (22-bit ints)
    h = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0 = h[y0] >> 4
    r1 = h[y1] >> 4
    out = (((b0[:, None, None] * r0) >> 16) + ((b1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# getPerspectiveTransform / warpPerspective  (transform.rs:222-234)
# ----------------------------------------------------------------------------
def perspective_system(src_pts, dst_pts):
    """The 8x8 system OpenCV builds: maps src (x,y) -> dst (u,v)."""
    a = np.zeros((8, 8), np.float64)
    b = np.zeros(8, np.float64)
    for i in range(4):
        x, y = float(src_pts[i][0]), float(src_pts[i][1])
        u, v = float(dst_pts[i][0]), float(dst_pts[i][1])
        a[i, 0], a[i, 1], a[i, 2] = x, y, 1.0
        a[i, 6], a[i, 7] = -x * u, -y * u
        a[i + 4, 3], a[i + 4, 4], a[i + 4, 5] = x, y, 1.0
        a[i + 4, 6], a[i + 4, 7] = -x * v, -y * v
        b[i], b[i + 4] = u, v
    return a, b


def get_perspective_transform_ge(src_pts, dst_pts) -> np.ndarray:
    """Same system solved by f64 Gaussian elimination with partial pivoting --
    what the CUDA ROI kernel does per ROI.  The reference passes INTER_LINEAR (==1
    == DECOMP_SVD) as the solve method (transform.rs:222); SVD and GE agree to
    ~1e-12 relative, far below the 1/32-px quantisation of the warp."""
    a, b = perspective_system(np.asarray(src_pts, np.float32), np.asarray(dst_pts, np.float32))
    a = a.copy()
    b = b.copy()
    n = 8
    for k in range(n):
        p = k + int(np.argmax(np.abs(a[k:, k])))
        if p != k:
            a[[k, p]] = a[[p, k]]
            b[[k, p]] = b[[p, k]]
        for i in range(k + 1, n):
            f = a[i, k] / a[k, k]
            a[i, k:] -= f * a[k, k:]
            b[i] -= f * b[k]
    x = np.zeros(n)
    for k in range(n - 1, -1, -1):
        x[k] = (b[k] - np.dot(a[k, k + 1:], x[k + 1:])) / a[k, k]
    return np.append(x, 1.0).reshape(3, 3)


def invert3x3(m: np.ndarray) -> np.ndarray:
    """cv::invert for a 3x3 f64 matrix (closed-form cofactors, as OpenCV does)."""
    m = np.asarray(m, np.float64)
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1])
         - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    d = 1.0 / d
    t = np.empty((3, 3), np.float64)
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def _bilinear_tab():
    """OpenCV's 32x32 table of 4 int16 bilinear weights (sum == 32768)."""
    tab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 4), np.int32)
    one = np.float32(1.0)
    scale = np.float32(1.0 / INTER_TAB_SIZE)
    for ay in range(INTER_TAB_SIZE):
        fy = np.float32(ay) * scale
        for ax in range(INTER_TAB_SIZE):
            fx = np.float32(ax) * scale
            w = np.array([(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx], np.float32)
            iw = np.rint(w * np.float32(INTER_REMAP_COEF_SCALE)).astype(np.int32)
            # OpenCV fixes the sum to exactly 32768 by adjusting the largest/smallest
            # weight; for the bilinear table every entry already sums to 32768.
            assert iw.sum() == INTER_REMAP_COEF_SCALE
            tab[ay, ax] = iw
    return tab


_BTAB = None


def warp_perspective_u8(src: np.ndarray, m: np.ndarray, dsize) -> np.ndarray:
    """cv2.warpPerspective(src, M, (w,h), INTER_LINEAR, BORDER_CONSTANT, 0), uint8 HxWxC."""
    global _BTAB
    if _BTAB is None:
        _BTAB = _bilinear_tab()
    w, h = int(dsize[0]), int(dsize[1])
    mi = invert3x3(m)
    sh, sw = src.shape[:2]
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    ww = mi[2, 0] * x + mi[2, 1] * y + mi[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        ww = np.where(ww != 0, INTER_TAB_SIZE / ww, 0.0)
    fx = np.clip((mi[0, 0] * x + mi[0, 1] * y + mi[0, 2]) * ww, -2147483648.0, 2147483647.0)
    fy = np.clip((mi[1, 0] * x + mi[1, 1] * y + mi[1, 2]) * ww, -2147483648.0, 2147483647.0)
    X = np.rint(fx).astype(np.int64)
    Y = np.rint(fy).astype(np.int64)
    sx = (X >> INTER_BITS).astype(np.int64)
    sy = (Y >> INTER_BITS).astype(np.int64)
    # OpenCV stores the integer source coordinate as saturated int16
    sx = np.clip(sx, -32768, 32767)
    sy = np.clip(sy, -32768, 32767)
    ax = (X & (INTER_TAB_SIZE - 1)).astype(np.int64)
    ay = (Y & (INTER_TAB_SIZE - 1)).astype(np.int64)
    wts = _BTAB[ay, ax]                                   # [h,w,4]
    s = src.astype(np.int32)
    acc = np.zeros((h, w, src.shape[2]), np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy = sy + dy
        xx = sx + dx
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = s[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)]
        v = np.where(ok[..., None], v, 0)
        acc += v * wts[..., k][..., None]
    out = (acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
    return np.clip(out, 0, 255).astype(np.uint8)


def copy_make_border_const0(src: np.ndarray, top, bottom, left, right) -> np.ndarray:
    return np.pad(src, ((top, bottom), (left, right), (0, 0)))


def flip_horizontal(src: np.ndarray) -> np.ndarray:
    return src[:, ::-1].copy()
