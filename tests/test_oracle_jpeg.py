"""Frame ingest oracle: oracle/jpeg_decode.py against cv2.imdecode + BGR2RGB, i.e. against the very calls
`convert_image_to_mat` makes (reference src/face_detection_lite/utils.rs:8-21).  Bit-exact is the bar (u8 output)."""
import glob
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(buf: bytes) -> np.ndarray:
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def test_reference_test_images_decode_bit_exact():
    from oracle import jpeg_decode
    files = sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg")))
    assert len(files) >= 3                       # man.jpg, russ_cox_1.jpg (restart interval 25), russ_cox_2.jpg (225 rows: partial MCU row)
    for f in files:
        buf = open(f, "rb").read()
        np.testing.assert_array_equal(jpeg_decode.convert_image_to_mat(buf), _ref(buf), err_msg=f)


@pytest.mark.parametrize("sampling", ["444", "422", "420"])
def test_sampling_sizes_qualities_restarts(sampling):
    from oracle import jpeg_decode
    fac = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
    img = cv2.imread(os.path.join(ROOT, "test_data", "man.jpg"))
    rng = np.random.default_rng(int(sampling))
    for (h, w) in ((360, 540), (97, 131), (8, 8), (1, 1), (17, 33), (250, 3), (3, 250), (16, 16), (15, 17), (9, 4), (9, 5), (2, 2)):
        src = img[:h, :w] if h > 16 and w > 16 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)     # photo crops and pure noise
        for q in (35, 90, 100):
            for rst in (0, 3):
                ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, fac,
                                                    cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                assert ok
                buf = enc.tobytes()
                np.testing.assert_array_equal(jpeg_decode.decode_jpeg_rgb(buf), _ref(buf), err_msg="%s %dx%d q%d rst%d" % (sampling, w, h, q, rst))


def test_greyscale_optimised_tables_and_rejections():
    from oracle import jpeg_decode
    img = cv2.imread(os.path.join(ROOT, "test_data", "russ_cox_1.jpg"))
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 80])
    np.testing.assert_array_equal(jpeg_decode.decode_jpeg_rgb(enc.tobytes()), _ref(enc.tobytes()))
    noise = np.random.default_rng(5).integers(0, 256, (123, 77, 3), dtype=np.uint8)
    ok, enc = cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_OPTIMIZE, 1])      # per-image Huffman tables
    np.testing.assert_array_equal(jpeg_decode.decode_jpeg_rgb(enc.tobytes()), _ref(enc.tobytes()))
    ok, enc = cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError):
        jpeg_decode.decode_jpeg_rgb(enc.tobytes())
    with pytest.raises(ValueError):
        jpeg_decode.decode_jpeg_rgb(b"\x89PNG\r\n")


def _host_backend(buf: bytes) -> np.ndarray:
    """csrc/jpeg_math.h on the host (tests/hostcheck): the oracle's entropy stage feeds quantised blocks to the per-block /
    per-pixel functions a device decoder will call."""
    import ctypes as C
    import hostcheck
    from oracle import jpeg_decode
    hc = hostcheck.load()
    f = jpeg_decode.entropy_decode(buf)
    n = len(f["comps"])
    coefs = [np.ascontiguousarray(c["coef"], np.int16) for c in f["comps"]]
    quants = [np.ascontiguousarray(c["quant"], np.uint16) for c in f["comps"]]
    for c, a in zip(f["comps"], coefs):
        assert (a == c["coef"]).all()                      # baseline coefficients fit 16 bits
    cp = (C.c_void_p * n)(*[a.ctypes.data for a in coefs])
    qp = (C.c_void_p * n)(*[a.ctypes.data for a in quants])
    samp = (C.c_int * (2 * n))(*[v for c in f["comps"] for v in (c["h"], c["v"])])
    out = np.empty((f["H"], f["W"], 3), np.uint8)
    hc.hc_jpeg_backend.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert hc.hc_jpeg_backend(n, cp, qp, samp, f["W"], f["H"], out.ctypes.data) == 0
    return out


def test_device_back_half_header_is_bit_exact_on_the_host():
    """jpeg_math.h (inverse DCT, gather-form fancy upsampling, colour conversion) == cv2.imdecode on the reference's images and
    on re-encoded variants of every supported sampling, including 1-pixel-wide / 1-pixel-high chroma planes."""
    files = sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg")))
    for f in files:
        buf = open(f, "rb").read()
        np.testing.assert_array_equal(_host_backend(buf), _ref(buf), err_msg=f)
    img = cv2.imread(os.path.join(ROOT, "test_data", "man.jpg"))
    rng = np.random.default_rng(9)
    for sampling in ("444", "422", "420"):
        fac = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
        for (h, w) in ((97, 131), (1, 1), (2, 2), (17, 33), (250, 3), (3, 250), (16, 16)):
            src = img[:h, :w] if h > 16 and w > 16 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            for q in (50, 100):
                ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, fac])
                np.testing.assert_array_equal(_host_backend(enc.tobytes()), _ref(enc.tobytes()), err_msg="%s %dx%d q%d" % (sampling, w, h, q))
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 75])
    np.testing.assert_array_equal(_host_backend(enc.tobytes()), _ref(enc.tobytes()))


def test_entropy_building_blocks_match_the_oracle():
    """csrc/jpeg_math.h's Huffman table / bit reader / block decoder (the sequential core of a device entropy stage), run on
    the host over whole scans, reproduce the oracle's quantised coefficients exactly -- restart intervals, optimised tables,
    4:4:4 / 4:2:2 / 4:2:0 / greyscale."""
    import ctypes as C
    import hostcheck
    from oracle import jpeg_decode
    hc = hostcheck.load()
    hc.hc_jpeg_entropy.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    bufs = [open(f, "rb").read() for f in sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg")))]
    img = cv2.imread(os.path.join(ROOT, "test_data", "man.jpg"))
    noise = np.random.default_rng(4).integers(0, 256, (61, 83, 3), dtype=np.uint8)
    for sampling in ("444", "422", "420"):
        fac = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
        for src, q, rst, opt in ((img[:200, :300], 85, 0, 0), (img[:97, :131], 40, 2, 1), (noise, 100, 5, 0), (noise, 60, 0, 1)):
            ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, fac, cv2.IMWRITE_JPEG_RST_INTERVAL, rst,
                                                cv2.IMWRITE_JPEG_OPTIMIZE, opt])
            bufs.append(enc.tobytes())
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 75])
    bufs.append(enc.tobytes())
    checked_intervals = []
    for buf in bufs:
        f = jpeg_decode.entropy_decode(buf)
        n = len(f["comps"])
        tabs = [np.frombuffer(t[0] + t[1], np.uint8).copy() for c in f["comps"] for t in (c["dht_dc"], c["dht_ac"])]
        tp = (C.c_void_p * (2 * n))(*[t.ctypes.data for t in tabs])
        outs = [np.full(c["coef"].shape, 12345, np.int16) for c in f["comps"]]
        op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        samp = (C.c_int * (2 * n))(*[v for c in f["comps"] for v in (c["h"], c["v"])])
        data = np.frombuffer(buf, np.uint8)
        assert hc.hc_jpeg_entropy(data.ctypes.data, len(buf), f["scan_offset"], n, samp, tp, f["restart_interval"], f["W"], f["H"], op) == 0
        for c, o in zip(f["comps"], outs):
            np.testing.assert_array_equal(o, c["coef"])
        if f["restart_interval"]:
            # the parallel schedule: every restart interval decoded independently (in reverse order), located by a byte scan
            for o in outs:
                o.fill(12345)
            mcus = (-(-f["W"] // (8 * f["hmax"]))) * (-(-f["H"] // (8 * f["vmax"])))
            hc.hc_jpeg_entropy_by_interval.argtypes = hc.hc_jpeg_entropy.argtypes
            n_iv = hc.hc_jpeg_entropy_by_interval(data.ctypes.data, len(buf), f["scan_offset"], n, samp, tp, f["restart_interval"], f["W"], f["H"], op)
            assert n_iv == -(-mcus // f["restart_interval"])
            for c, o in zip(f["comps"], outs):
                np.testing.assert_array_equal(o, c["coef"])
            checked_intervals.append(n_iv)
    assert len(checked_intervals) >= 7 and max(checked_intervals) > 20


def test_host_decoder_end_to_end_and_rejections():
    """jpeg_parse.h + jpeg_math.h chained on the host (parse -> entropy -> back half): bytes in, cv2.imdecode's pixels out;
    what the oracle rejects, the parser rejects too."""
    import ctypes as C
    import hostcheck
    from oracle import jpeg_decode
    hc = hostcheck.load()
    hc.hc_jpeg_decode.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]

    def decode(buf):
        data = np.frombuffer(buf, np.uint8)
        w, h = C.c_int(), C.c_int()
        err = C.create_string_buffer(128)
        if hc.hc_jpeg_decode(data.ctypes.data, len(buf), None, C.byref(w), C.byref(h), err, 128) != 0:
            raise ValueError(err.value.decode())
        out = np.empty((h.value, w.value, 3), np.uint8)
        assert hc.hc_jpeg_decode(data.ctypes.data, len(buf), out.ctypes.data, C.byref(w), C.byref(h), err, 128) == 0
        return out

    for f in sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg"))):
        buf = open(f, "rb").read()
        np.testing.assert_array_equal(decode(buf), _ref(buf), err_msg=f)
    frame = np.random.default_rng(2).integers(0, 256, (270, 480, 3), dtype=np.uint8)
    frame[60:200, 100:380] = cv2.resize(cv2.imread(os.path.join(ROOT, "test_data", "man.jpg")), (280, 140))
    for fac in ("444", "422", "420"):
        ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + fac),
                                               cv2.IMWRITE_JPEG_RST_INTERVAL, 7])
        np.testing.assert_array_equal(decode(enc.tobytes()), _ref(enc.tobytes()))
    ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    for bad in (enc.tobytes(), b"\x89PNG\r\n\x1a\n", open(os.path.join(ROOT, "test_data", "man.jpg"), "rb").read()[:300]):
        with pytest.raises(ValueError) as e_c:
            decode(bad)
        with pytest.raises(ValueError):
            jpeg_decode.decode_jpeg_rgb(bad)
        assert str(e_c.value)
    with pytest.raises(ValueError, match="baseline"):
        decode(enc.tobytes())


def test_self_synchronising_schedule_matches_sequential_decode():
    """Files without restart markers (two of the reference's three test images): the scan is cut into fixed windows that are
    decoded from guessed states and re-decoded until the hand-over of exit states reaches a fixed point (csrc/jpeg_math.h
    jpeg_sync_step; the schedule itself runs on the host here).  The result must equal the sequential decode, and the fixed point
    must come after a few rounds -- the property that makes the scheme parallel."""
    import ctypes as C
    import hostcheck
    from oracle import jpeg_decode
    hc = hostcheck.load()
    hc.hc_jpeg_entropy_selfsync.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
    bufs = [open(os.path.join(ROOT, "test_data", n), "rb").read() for n in ("man.jpg", "russ_cox_2.jpg")]
    img = cv2.imread(os.path.join(ROOT, "test_data", "man.jpg"))
    noise = np.random.default_rng(8).integers(0, 256, (120, 200, 3), dtype=np.uint8)
    for sampling in ("444", "422", "420"):
        fac = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
        for src, q, opt in ((img, 90, 0), (img[:97, :131], 30, 1), (noise, 95, 0)):
            ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, fac, cv2.IMWRITE_JPEG_OPTIMIZE, opt])
            bufs.append(enc.tobytes())
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 75])
    bufs.append(enc.tobytes())
    worst = 0
    for buf in bufs:
        f = jpeg_decode.entropy_decode(buf)
        assert f["restart_interval"] == 0
        n = len(f["comps"])
        tabs = [np.frombuffer(t[0] + t[1], np.uint8).copy() for c in f["comps"] for t in (c["dht_dc"], c["dht_ac"])]
        tp = (C.c_void_p * (2 * n))(*[t.ctypes.data for t in tabs])
        samp = (C.c_int * (2 * n))(*[v for c in f["comps"] for v in (c["h"], c["v"])])
        data = np.frombuffer(buf, np.uint8)
        for window_bits in (1024, 256):
            outs = [np.full(c["coef"].shape, 12345, np.int16) for c in f["comps"]]
            op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
            nwin, redone = C.c_int(), C.c_int()
            rounds = hc.hc_jpeg_entropy_selfsync(data.ctypes.data, len(buf), f["scan_offset"], n, samp, tp, f["W"], f["H"], window_bits, op,
                                                 C.byref(nwin), C.byref(redone))
            assert rounds >= 1, rounds
            for c, o in zip(f["comps"], outs):
                np.testing.assert_array_equal(o, c["coef"])
            worst = max(worst, rounds)
            assert rounds <= 12 or rounds < nwin.value // 4, (rounds, nwin.value, redone.value)     # nowhere near sequential
    assert worst >= 2          # the guesses were wrong somewhere: the hand-over was exercised


def test_malformed_tables_and_damaged_scans_under_asan_and_ubsan(tmp_path):
    """a DHT whose code lengths are over-subscribed used to write past the 512-entry look-up table, and a DC
    symbol > 16 shifted by a negative amount.  tests/hostcheck/jpeg_fuzz.cc drives csrc/jpeg_parse.h + csrc/jpeg_math.h (parser,
    table build, sequential and window decoders, IDCT) over > 1500 damaged copies of the reference's test images -- targeted DHT
    damage, random header / scan damage, truncated scans -- compiled with -fsanitize=address,undefined: any out-of-bounds access,
    signed overflow or bad shift aborts the child process."""
    import subprocess
    exe = str(tmp_path / "jpeg_fuzz")
    src = os.path.join(ROOT, "tests", "hostcheck", "jpeg_fuzz.cc")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", src, "-o", exe])
    files = sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg")))
    r = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "no sanitizer report" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
    # and the oracle agrees with the parser on the textbook case: three 1-bit codes
    from oracle import jpeg_decode
    buf = bytearray(open(files[0], "rb").read())
    p = buf.index(b"\xff\xc4")
    buf[p + 5] = 3
    with pytest.raises(ValueError):
        jpeg_decode.decode_jpeg_rgb(bytes(buf))


def test_colour_kernel_regrouping_is_exact():
    """jpeg_color.cuh evaluates jdsample.c / jdcolor.c with the operations regrouped (byte dot products whose rounding constant
    carries -128, `((y << 16) + ONE_HALF + k * x) >> 16`, a saturating pack as the clamp).  The regrouped integer expressions,
    restated here in numpy, must equal the oracle's straight forms: every (y, cb, cr) for the colour step, random and extreme
    planes of odd and even sizes for both upsamplers."""
    from oracle import jpeg_decode as J

    # ---- YCbCr -> RGB, all 2^24 inputs (16 luma values at a time)
    for y0 in range(0, 256, 16):
        y, cb, cr = np.meshgrid(np.arange(y0, y0 + 16, dtype=np.int64), np.arange(256, dtype=np.int64), np.arange(256, dtype=np.int64), indexing="ij")
        xb, xr = cb - 128, cr - 128
        yh = (y << 16) + 32768
        r = (91881 * xr + yh) >> 16
        g = (-22554 * xb - 46802 * xr + yh) >> 16
        b = (116130 * xb + yh) >> 16
        got = np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)
        assert np.array_equal(got, J._ycc_to_rgb(y.astype(np.uint8), cb, cr))

    # ---- fancy upsampling as weighted sums with the centring folded into the rounding constant
    def up_h2v2(c):
        h, w = c.shape
        ci = c.astype(np.int64)
        out = np.empty((2 * h, 2 * w), np.int64)
        for Y in range(2 * h):
            cy = Y >> 1
            fy = min(cy + 1, h - 1) if Y & 1 else max(cy - 1, 0)
            n, f = ci[cy], ci[fy]
            left_n, left_f = np.roll(n, 1), np.roll(f, 1)
            right_n, right_f = np.roll(n, -1), np.roll(f, -1)
            even = (9 * n + 3 * f + 3 * left_n + left_f + (8 - 2048)) >> 4          # weights 3x(3,1) near, (3,1) far
            odd = (9 * n + 3 * f + 3 * right_n + right_f + (7 - 2048)) >> 4
            even[0] = (12 * n[0] + 4 * f[0] + (8 - 2048)) >> 4                       # edges: weight 4 on the sample itself
            odd[-1] = (12 * n[-1] + 4 * f[-1] + (7 - 2048)) >> 4
            out[Y, 0::2] = even + 128
            out[Y, 1::2] = odd + 128
        return out.astype(np.uint8)

    def up_h2v1(c):
        ci = c.astype(np.int64)
        left, right = np.roll(ci, 1, axis=1), np.roll(ci, -1, axis=1)
        even = (3 * ci + left + (1 - 512)) >> 2
        odd = (3 * ci + right + (2 - 512)) >> 2
        even[:, 0] = (4 * ci[:, 0] + (1 - 512)) >> 2
        odd[:, -1] = (4 * ci[:, -1] + (2 - 512)) >> 2
        out = np.empty((c.shape[0], 2 * c.shape[1]), np.int64)
        out[:, 0::2] = even + 128
        out[:, 1::2] = odd + 128
        return out.astype(np.uint8)

    rng = np.random.default_rng(5)
    for h, w in ((1, 3), (2, 3), (5, 4), (8, 17), (33, 64)):
        for plane in (rng.integers(0, 256, (h, w), dtype=np.uint8), np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8),
                      (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)):
            assert np.array_equal(up_h2v2(plane), J._upsample_h2v2(plane))
            assert np.array_equal(up_h2v1(plane), J._upsample_h2v1(plane))
