"""Frame ingest on the device: fdl_decode_jpeg / fdl_jpeg_decode / fdl_pipeline_submit_jpeg against the very
calls `convert_image_to_mat` makes (reference src/face_detection_lite/utils.rs:8-21: cv2.imdecode(IMREAD_COLOR) + BGR2RGB).
u8 output: bit-exact is the bar.  The same files (the reference's three test images and the 216 re-encoded variants of
tests/test_oracle_jpeg.py) pinned the oracle and the host headers on the CPU; here they go through the kernels."""
import glob
import os

import numpy as np
import pytest

from conftest import MODELS

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ref(buf: bytes) -> np.ndarray:
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(buf, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def _variants(sampling):
    fac = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
    img = cv2.imread(os.path.join(ROOT, "test_data", "man.jpg"))
    rng = np.random.default_rng(int(sampling))
    out = []
    for (h, w) in ((360, 540), (97, 131), (8, 8), (1, 1), (17, 33), (250, 3), (3, 250), (16, 16), (15, 17), (9, 4), (9, 5), (2, 2)):
        src = img[:h, :w] if h > 16 and w > 16 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for q in (35, 90, 100):
            for rst in (0, 3):
                ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, fac, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                assert ok
                out.append(("%s %dx%d q%d rst%d" % (sampling, w, h, q, rst), enc.tobytes()))
    return out


def test_reference_test_images_decode_bit_exact(fdl, gpu):
    files = sorted(glob.glob(os.path.join(ROOT, "test_data", "*.jpg")))
    assert len(files) >= 3          # man.jpg, russ_cox_1.jpg (restart interval 25), russ_cox_2.jpg (225 rows: partial MCU row)
    for f in files:
        buf = open(f, "rb").read()
        assert fdl.jpeg_info(buf)[:2] == _ref(buf).shape[1::-1]
        np.testing.assert_array_equal(fdl.convert_image_to_mat(buf, device=gpu), _ref(buf), err_msg=f)


@pytest.mark.parametrize("sampling", ["444", "422", "420"])
def test_sampling_sizes_qualities_restarts(fdl, gpu, sampling):
    """72 files per sampling mode -- sizes down to 1x1, qualities 35..100, with and without restart intervals -- decoded as ONE
    mixed batch (different sizes, tables and entropy schedules side by side in the same launches), every pixel against cv2."""
    dec = fdl.JpegDecoder(device=gpu)
    cases = _variants(sampling)
    got = dec.decode([b for _, b in cases])
    assert len(got) == len(cases)
    for (name, buf), g in zip(cases, got):
        np.testing.assert_array_equal(g, _ref(buf), err_msg=name)
    # and one at a time, in reverse order (stale work buffers of a larger batch must not leak into a smaller one)
    for name, buf in cases[::-7]:
        np.testing.assert_array_equal(dec.decode([buf])[0], _ref(buf), err_msg=name)
    dec.close()


def test_greyscale_optimised_tables_large_and_rejections(fdl, gpu):
    dec = fdl.JpegDecoder(device=gpu)
    img = cv2.imread(os.path.join(ROOT, "test_data", "russ_cox_1.jpg"))
    bufs = []
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 80]); bufs.append(enc.tobytes())
    noise = np.random.default_rng(5).integers(0, 256, (123, 77, 3), dtype=np.uint8)
    ok, enc = cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_OPTIMIZE, 1]); bufs.append(enc.tobytes())   # per-image Huffman tables
    # a 4K noise image at quality 100: ~20 MB of scan, far more than 4096 windows of 1024 bits -> the window size grows instead
    big = np.random.default_rng(6).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
    ok, enc = cv2.imencode(".jpg", big, [cv2.IMWRITE_JPEG_QUALITY, 100]); bufs.append(enc.tobytes())
    ok, enc = cv2.imencode(".jpg", big[:1080, :1920], [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_RST_INTERVAL, 120]); bufs.append(enc.tobytes())
    for b, g in zip(bufs, dec.decode(bufs)):
        np.testing.assert_array_equal(g, _ref(b))
    # device-resident output
    t, offs, ws, hs = dec.decode_to_device(bufs[:2])
    host = t.cpu().numpy()
    for i in range(2):
        np.testing.assert_array_equal(host[offs[i]:offs[i] + ws[i] * hs[i] * 3].reshape(hs[i], ws[i], 3), _ref(bufs[i]))
    ok, prog = cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    man = open(os.path.join(ROOT, "test_data", "man.jpg"), "rb").read()
    for bad, needle in ((prog.tobytes(), "baseline"), (b"\x89PNG\r\n\x1a\n", "JPEG"), (man[:300], ""), (man[:len(man) // 2], "premature end")):
        with pytest.raises(fdl.FdlError) as e:
            dec.decode([bufs[1], bad])
        assert e.value.code == -1 and needle in e.value.message, e.value.message
    np.testing.assert_array_equal(dec.decode([bufs[1]])[0], _ref(bufs[1]))      # the handle survives a rejected batch
    dec.close()


def test_pipeline_from_jpeg_bytes_equals_pipeline_from_decoded_frames(fdl, gpu):
    """lib.rs:20-40 from where it starts: encoded bytes.  Pipeline.submit_jpeg (compressed H2D + device decode) must give exactly
    the results of Pipeline.submit on the frames cv2 decodes from the same bytes -- plain list of files, and the pinned-arena form."""
    import synth_frames
    import torch
    n = 6
    frames = synth_frames.face_frames(n, start=40, faces=("man.jpg", "russ_cox_1.jpg", "russ_cox_2.jpg"))
    files = []
    for i in range(n):
        ok, enc = cv2.imencode(".jpg", frames[i][:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, 120] if i == 2 else []))
        files.append(enc.tobytes())
    decoded = np.stack([_ref(b) for b in files])
    pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=8, max_faces=1, model_dir=MODELS, device=gpu)
    want = pipe.run(decoded)
    assert sum(1 for r in want for f in r.faces if f.landmarks is not None) >= n - 1
    offs = np.cumsum([0] + [(len(b) + 63) & ~63 for b in files])
    arena = torch.zeros(int(offs[-1]), dtype=torch.uint8).pin_memory()
    for i, b in enumerate(files):
        arena[offs[i]:offs[i] + len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
    t1 = pipe.submit_jpeg(files[:4])
    t2 = pipe.submit_jpeg(files[4:])
    t3 = pipe.submit_jpeg((arena, offs[:-1], [len(b) for b in files]))
    for got in (pipe.collect(t1) + pipe.collect(t2), pipe.collect(t3)):
        assert len(got) == n
        for a, b in zip(want, got):
            assert [d.anchor for d in a.detections] == [d.anchor for d in b.detections]
            for da, db in zip(a.detections, b.detections):
                np.testing.assert_array_equal(da.data, db.data)
            for fa, fb in zip(a.faces, b.faces):
                assert (fa.landmarks is None) == (fb.landmarks is None)
                if fa.landmarks is not None:
                    np.testing.assert_array_equal(fa.landmarks, fb.landmarks)
                    np.testing.assert_array_equal(fa.left_iris, fb.left_iris)
                    np.testing.assert_array_equal(fa.right_contour, fb.right_contour)
    with pytest.raises(fdl.FdlError):
        pipe.run_jpeg([open(os.path.join(ROOT, "test_data", "man.jpg"), "rb").read()])       # 540x360 into a 1080p pipeline
    with pytest.raises(fdl.FdlError) as e:
        pipe.run_jpeg([files[0], files[1][:len(files[1]) // 2]])                              # truncated scan: reported by collect
    assert "premature end" in e.value.message
    assert len(pipe.run_jpeg(files[:2])) == 2                                                 # and the lane is free again
    pipe.close()


def test_sparse_colour_conversion_reads_no_unconverted_pixel(gpu):
    """Pipeline.submit_jpeg converts only the letterbox rows and the ROI spans of the decoded frames (pipeline.cu, jpeg_color_rows /
    jpeg_color_roi kernels).  With the frame buffer poisoned before every decode (FDL_JPEG_POISON=1, a fresh process: the switch is
    read once) the results must still equal the pipeline's on the frames cv2 decodes: one and two faces per frame, 4:2:0 / 4:2:2,
    restart markers, two frame sizes, a batch off the colour fast path (4:4:4: whole frames), the short-range detector alone."""
    import subprocess
    import sys
    env = dict(os.environ, FDL_JPEG_POISON="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "jpeg_sparse_check.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sparse ok" in r.stdout
