"""Host logic without a GPU: the C ABI loads and exports every symbol of include/fdl.h, the C++ flatbuffer
reader + planner agree with the oracle's independent reader, the glue arithmetic compiled for the host matches
the oracle, calls that need a device fail loudly, and the rank-sharding used by bench.py works under gloo."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import MODELS, ROOT, rng

MODEL_FILES = ["face_detection_short_range", "face_detection_front", "face_detection_back", "face_detection_full_range", "face_landmark",
               "iris_landmark", "face_detection_full_range_sparse"]


def test_library_exports_every_declared_symbol(fdl):
    from rs_face_detection_tflite_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fdl.h")).read()
    declared = set(re.findall(r"FDL_API\s+[^;{]*?\b(fdl_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (fdl_\w+)", out))
    assert declared <= exported
    assert lib.fdl_version().startswith(b"fdl-b200")


def test_struct_layouts_match_header(fdl):
    from rs_face_detection_tflite_b200 import _lib
    assert C.sizeof(_lib.CRect) == 48 and C.sizeof(_lib.CDetection) == 72 and C.sizeof(_lib.CLandmark) == 24
    assert C.sizeof(_lib.CImage) == 32
    assert C.sizeof(_lib.CFrameResult) == 12 + 32 * 72
    assert C.sizeof(_lib.CFaceResult) == 48 + 8 + 468 * 12 + 96 + 2 * 71 * 12 + 2 * 5 * 12 + 468 * 12 + 16 + 16
    assert C.sizeof(_lib.CPipelineConfig) == 8 * 4 + 8 + 2 * 4 + 8


@pytest.mark.skipif("torch.cuda.is_available()", reason="checks the no-GPU error path")
def test_no_device_is_an_error_not_a_fallback(fdl):
    import torch  # noqa: F401
    assert fdl.device_count() == 0
    for make in (lambda: fdl.FaceDetection(fdl.FaceDetectionModel.BackCamera, MODELS), lambda: fdl.FaceLandmark(MODELS + "/face_landmark.tflite"),
                 lambda: fdl.IrisLandmark(MODELS + "/iris_landmark.tflite"), lambda: fdl.Pipeline(model_dir=MODELS),
                 lambda: fdl.face_detection_to_roi(fdl.Detection(np.zeros((8, 2), np.float32), 0.9), (10, 10)),
                 lambda: fdl.image_to_tensor(np.zeros((8, 8, 3), np.uint8), None, (4, 4), True),
                 lambda: fdl.Pool([0], model_dir=MODELS), lambda: fdl.Frame(np.zeros((8, 8, 3), np.uint8)),
                 lambda: fdl.convert_image_to_mat(open(os.path.join(ROOT, "test_data", "man.jpg"), "rb").read())):
        with pytest.raises(fdl.FdlError) as e:
            make()
        assert e.value.code == -4 and "no CPU fallback" in e.value.message


try:
    import torch  # noqa: F401
except Exception:  # pragma: no cover
    torch = None


@pytest.mark.parametrize("name", MODEL_FILES)
def test_planner_covers_every_op_and_matches_oracle_reader(fdl, name):
    """Plan-only handle (device -1): every non-constant tflite op lands in exactly one launch; shapes, launch counts and
    FLOPs agree with what the oracle's independent flatbuffer reader sees."""
    from oracle import tflite_reader as T
    path = os.path.join(MODELS, name + ".tflite")
    net = fdl.Net(path, device=-1)
    text = net.describe()
    m = T.load(path)
    head = re.search(r"plan: (\d+) tflite ops -> (\d+) launches; arena (\d+) floats/item; weights (\d+) floats; block-fused floor (\d+) bytes/item; (\d+) flop/item", text)
    n_ops, n_launch, arena, weights, floor, flops = map(int, head.groups())
    assert n_ops == len(m.ops) and n_launch == net.num_steps
    covered = []
    for line in text.splitlines()[1:]:
        if not line.startswith("#"):
            continue                                                   # (the tail chain's program follows the step list)
        covered += [int(v) for v in re.search(r"ops=\[([\d,]*)\]", line).group(1).split(",") if v]
    folded = {T.DEQUANTIZE, T.RESHAPE, T.CONCATENATION, T.DENSIFY}
    expect = [i for i, op in enumerate(m.ops) if op.code not in folded]
    assert sorted(covered) == expect
    # FLOPs of conv + depthwise ops from the oracle reader
    ref_flops = 0
    for op in m.ops:
        if op.code in (T.CONV_2D, T.DEPTHWISE_CONV_2D):
            w = m.tensors[op.inputs[1]].shape
            o = m.tensors[op.outputs[0]].shape
            k = w[1] * w[2] * (w[3] if op.code == T.CONV_2D else 1)
            ref_flops += 2 * k * o[1] * o[2] * o[3]
    assert flops == ref_flops
    assert floor > 0 and arena > 0
    with pytest.raises(fdl.FdlError):
        net.forward(np.zeros([1] + m.tensors[m.inputs[0]].shape[1:], np.float32))   # plan-only handles cannot run
    net.close()


def test_planner_fuses_blaze_blocks(fdl):
    net = fdl.Net(os.path.join(MODELS, "face_detection_back.tflite"), device=-1)
    text = net.describe()
    assert net.num_steps == 37                       # 282 tflite ops
    assert text.count("BLOCK dw3x3/s1+pw 24->24 in 128x128") == 7
    assert "skip=maxpool+chanpad" in text and "CONV 5x5/s2 3->24" in text
    assert "block-fused floor 37293568 bytes/item" in text and "377499648 flop/item" in text
    net.close()


def test_bad_models_are_rejected(fdl, tmp_path):
    with pytest.raises(fdl.FdlError) as e:
        fdl.Net(str(tmp_path / "missing.tflite"), device=-1)
    assert e.value.code == -2
    bad = tmp_path / "bad.tflite"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(fdl.FdlError) as e:
        fdl.Net(str(bad), device=-1)
    assert e.value.code == -3
    # a truncated real file must be rejected by the bounds checks, not crash
    data = open(os.path.join(MODELS, "face_detection_back.tflite"), "rb").read()
    cut = tmp_path / "cut.tflite"
    cut.write_bytes(data[: len(data) // 3])
    with pytest.raises(fdl.FdlError):
        fdl.Net(str(cut), device=-1)


# ---- glue arithmetic (csrc/glue_math.h, host build) against the oracle --------------------------------
@pytest.fixture(scope="module")
def hc():
    import hostcheck
    return hostcheck.load()


def test_host_anchors_bit_exact(hc):
    from oracle import glue
    for model in (0, 1, 2, 3):
        ref = glue.ssd_generate_anchors(model)
        out = np.empty_like(ref)
        assert hc.hc_anchors(model, out.ctypes.data_as(C.c_void_p), len(ref)) == len(ref)
        np.testing.assert_array_equal(out, ref)


def test_host_postprocess_matches_oracle(hc, fdl):
    from oracle import glue
    from rs_face_detection_tflite_b200._lib import CDetection
    r = rng(77)
    for model, n, size in ((1, 896, 256), (3, 2304, 192)):
        anchors = glue.ssd_generate_anchors(model)
        for case in range(8):
            reg = r.normal(0, 20, (n, 16)).astype(np.float32)
            cls = r.normal(-6, 2, (n, 1)).astype(np.float32)
            for _ in range(case % 4 + 1):
                c = int(r.integers(0, n))
                base = r.normal(0, 10, 16).astype(np.float32)
                base[2:4] = r.uniform(0.15, 0.5, 2) * size
                for j in range(int(r.integers(1, 7))):
                    k = (c + int(r.integers(-3, 4))) % n
                    reg[k] = base + r.normal(0, 2, 16)
                    cls[k, 0] = r.uniform(0.2, 6)
            pad = (0.0, 0.21875, 0.0, 0.21875) if case % 2 else (0.0, 0.0, 0.0, 0.0)
            dets = glue.convert_to_detections(glue.decode_boxes(reg, anchors, float(size)), glue.get_sigmoid_score(cls))
            ref = glue.detection_letterbox_removal(glue.non_maximum_suppression(dets), pad)
            out = (CDetection * 64)()
            surv = (C.c_int * n)()
            ns = C.c_int()
            cnt = hc.hc_ssd_postprocess(model, reg.ctypes.data_as(C.c_void_p), cls.ctypes.data_as(C.c_void_p), (C.c_double * 4)(*pad), out, 64, surv,
                                        C.byref(ns))
            assert list(surv[:ns.value]) == [d.anchor for d in dets]
            assert cnt == len(ref)
            for k, e in enumerate(ref):
                assert out[k].anchor == e.anchor and np.float32(out[k].score) == e.score
                np.testing.assert_allclose(np.array(out[k].data[:], np.float32).reshape(8, 2), e.data, atol=1e-6, rtol=0)


def test_host_roi_and_projection_match_oracle(hc):
    from oracle import glue
    from rs_face_detection_tflite_b200._lib import CRect
    r = rng(5)
    for k in range(20):
        data = r.uniform(0.1, 0.9, (8, 2)).astype(np.float32)
        data[1] = data[0] + r.uniform(0.05, 0.3, 2).astype(np.float32)
        size = (int(r.integers(100, 2000)), int(r.integers(100, 2000)))
        ref = glue.face_detection_to_roi(glue.Detection(data, np.float32(0.9)), size)
        out = CRect()
        assert hc.hc_face_detection_to_roi(data.ctypes.data_as(C.c_void_p), size[0], size[1], -1, C.byref(out)) == 0
        for f in ("x_center", "y_center", "width", "height", "rotation"):
            assert abs(getattr(out, f) - getattr(ref, f)) <= 1e-13 * max(1.0, abs(getattr(ref, f)))
        raw = r.uniform(-20, 220, (71, 3)).astype(np.float32)
        roi = glue.Rect(r.uniform(0.2, 0.8), r.uniform(0.2, 0.8), r.uniform(0.1, 0.9), r.uniform(0.1, 0.9), r.uniform(-3, 3), True)
        pad = (1.1e-16, 0.0, 1.1e-16, 0.0) if k % 2 else (0.0, 0.0, 0.0, 0.0)
        refp = glue.project_landmarks(raw, (64, 64), size, pad, roi, bool(k & 2))
        outp = np.empty((71, 3), np.float32)
        croi = CRect(roi.x_center, roi.y_center, roi.width, roi.height, roi.rotation, 1, 0)
        hc.hc_project(raw.ctypes.data_as(C.c_void_p), 71, 64, 64, size[0], size[1], (C.c_double * 4)(*pad), C.byref(croi), int(bool(k & 2)),
                      outp.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(outp, refp, atol=2e-7, rtol=0)


# ---- multi-process plumbing (gloo, world size 2) -------------------------------------------------------
_WORKER = r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import synth_frames
# frames are sharded by rank exactly as bench.py does (start = rank * uniq); no data-path collective
uniq = 2
mine = synth_frames.face_frames(uniq, 320, 180, start=rank * uniq)
digest = torch.tensor([int(mine.astype("int64").sum())], dtype=torch.int64)
gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, digest)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max-over-ranks timing, as in bench.py
dist.barrier()
if rank == 0:
    print(json.dumps({"digests": [int(g) for g in gathered], "tmax": float(t)}))
dist.destroy_process_group()
"""


def test_rank_sharding_under_gloo(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert len(res["digests"]) == 2 and res["digests"][0] != res["digests"][1]   # ranks work on different frames
    assert res["tmax"] == 2.0


@pytest.mark.parametrize("size,s", [((1920, 1080), 256), ((1280, 720), 256), ((3840, 2160), 256), ((1920, 1080), 128), ((2560, 1440), 192)])
def test_letterbox_row_plan_covers_exactly_the_rows_opencv_reads(fdl, size, s):
    """Zero-copy ingest (pipeline.cu plan_row_gather): the copy engine must gather every source row the letterbox resize
    interpolates between -- checked against the oracle's restatement of OpenCV's INTER_LINEAR row mapping (cv_ops.py, B.1)
    -- and nothing else, and the pattern must be periodic across contiguous frames."""
    from oracle import cv_ops
    w, h = size
    plan = fdl.letterbox_row_plan(size, s)
    assert plan is not None, "16:9 frames have a periodic row pattern"
    row_pos, info = plan
    # transform.rs:239-280 for roi = None: bordered square of side w (pad_v rows above and below), one resize to s
    pad_v = int((1.0 - (h / w) / 1.0) / 2.0 * w)
    y0, _, _ = cv_ops._resize_axis_coeffs(s, w, clamp_frac=False)
    need = set()
    for y in y0:
        for r in (min(max(int(y), 0), w - 1), min(max(int(y) + 1, 0), w - 1)):
            if 0 <= r - pad_v < h:
                need.add(r - pad_v)
    got = {r for r in range(h) if row_pos[r] >= 0}
    assert got == need
    # compact indices are a bijection onto 0..rows_per_frame-1, ascending with the source row
    idx = [int(row_pos[r]) for r in sorted(got)]
    assert idx == list(range(info["rows_per_frame"]))
    assert info["period_src_rows"] * info["periods_per_frame"] == h      # continues seamlessly into the next frame
    assert len(need) < 0.6 * h and 1 <= info["copies"] <= 8


def test_letterbox_row_plan_declines_when_nothing_is_gained(fdl):
    assert fdl.letterbox_row_plan((256, 256), 256) is None        # no resize at all
    assert fdl.letterbox_row_plan((300, 200), 256) is None        # nearly every row is read


# ---- iris refinement: index maps, diameter, depth -------------------------------------------------
def test_eye_to_face_landmark_index_tables(fdl):
    """The oracle's tables, the library's host copy and the header the kernels use agree; when the reference tree is
    present (this container, not the GPU box) they are also checked against iris_landmark.rs:64-95 itself."""
    from oracle import glue
    import hostcheck
    hc = hostcheck.load()
    for eye, tab in ((0, glue.LEFT_EYE_TO_FACE_LANDMARK_INDEX), (1, glue.RIGHT_EYE_TO_FACE_LANDMARK_INDEX)):
        out = (C.c_int * 71)()
        assert hc.hc_eye_index(eye, out) == 71
        np.testing.assert_array_equal(np.array(out[:]), tab)
        np.testing.assert_array_equal(fdl.eye_to_face_landmark_index(bool(eye)), tab)
    assert not set(glue.LEFT_EYE_TO_FACE_LANDMARK_INDEX) & set(glue.RIGHT_EYE_TO_FACE_LANDMARK_INDEX)
    ref = "/root/reference/src/face_detection_lite/iris_landmark.rs"
    if os.path.exists(ref):
        src = open(ref).read()
        for name, tab in (("LEFT_EYE_TO_FACE_LANDMARK_INDEX", glue.LEFT_EYE_TO_FACE_LANDMARK_INDEX),
                          ("RIGHT_EYE_TO_FACE_LANDMARK_INDEX", glue.RIGHT_EYE_TO_FACE_LANDMARK_INDEX)):
            body = re.search(name + r": \[i32; 71\] = \[(.*?)\];", src, re.S).group(1)
            vals = [int(v) for v in re.findall(r"\d+", re.sub(r"//[^\n]*", "", body))]
            np.testing.assert_array_equal(np.array(vals), tab)


def test_oracle_update_face_landmarks_with_iris_results():
    from oracle import glue
    r = rng(5)
    face, left, right = r.random((468, 3)), r.random((71, 3)) + 2, r.random((71, 3)) + 4
    out = glue.update_face_landmarks_with_iris_results(face, left, right)
    li, ri = glue.LEFT_EYE_TO_FACE_LANDMARK_INDEX, glue.RIGHT_EYE_TO_FACE_LANDMARK_INDEX
    np.testing.assert_array_equal(out[li], left)
    np.testing.assert_array_equal(out[ri], right)
    rest = np.setdiff1d(np.arange(468), np.concatenate([li, ri]))
    assert len(rest) == 468 - 142
    np.testing.assert_array_equal(out[rest], face[rest])
    with pytest.raises(ValueError):
        glue.update_face_landmarks_with_iris_results(face[:100], left, right)


def test_iris_metrics_host_math_matches_oracle():
    """iris_diameter / iris_depth of csrc/glue_math.h (compiled for the host) == the oracle's f64 restatement, bit for bit,
    including the integer image centre of iris_landmark.rs:426 on odd sizes."""
    from oracle import glue
    import hostcheck
    hc = hostcheck.load()
    r = rng(11)
    for (w, h) in ((540, 360), (1920, 1080), (641, 479)):
        for _ in range(20):
            iris = r.random((5, 3)).astype(np.float32)
            d = glue.get_iris_diameter(iris.astype(np.float64), (w, h))
            z = glue.get_iris_depth(iris.astype(np.float64), 4.3, d, (w, h))
            out = (C.c_double * 2)()
            hc.hc_iris_metrics_f32(iris.ctypes.data_as(C.POINTER(C.c_float)), w, h, C.c_double(4.3), out)
            assert (out[0], out[1]) == (d, z)
            i64 = np.ascontiguousarray(iris, np.float64)
            hc.hc_iris_metrics(i64.ctypes.data_as(C.POINTER(C.c_double)), w, h, C.c_double(4.3), out)
            assert (out[0], out[1]) == (d, z)
    # closed form: a 10 px wide, 6 px high iris centred on the image centre, focal length f -> depth = 11.8 * f / 8
    w, h = 200, 100
    iris = np.array([[0.5, 0.5, 0], [0.475, 0.5, 0], [0.5, 0.47, 0], [0.525, 0.5, 0], [0.5, 0.53, 0]])
    assert abs(glue.get_iris_diameter(iris, (w, h)) - 8.0) < 1e-12
    assert abs(glue.get_iris_depth(iris, 5.0, 8.0, (w, h)) - 11.8 * 5.0 / 8.0) < 1e-12


# ---- sparse full-range detector: DENSIFY, spatial PAD, fused RELU, DEPTH_TO_SPACE ------------------
def test_sparse_model_densify_and_plan(fdl):
    """The oracle's CSR decode of the 46 sparse f16 weight tensors (277 552 B stored) gives dense tensors of the dense
    model's shapes with the stored number of non-zeros; the C++ planner folds DENSIFY / spatial PAD / fused RELU and
    plans DEPTH_TO_SPACE; both detectors' heads end up as branch streams."""
    from oracle import tflite_reader as T
    m = T.load(os.path.join(MODELS, "face_detection_full_range_sparse.tflite"))
    sparse = [t for t in m.tensors if t.sparsity is not None]
    assert len(sparse) == 46 and all(t.dtype == np.float16 for t in sparse)
    stored = 0
    for t in sparse:
        d = T.densify(t)
        assert list(d.shape) == t.shape and d.dtype == np.float16
        nnz = len(t.sparsity["dims"][-1][3])
        assert np.count_nonzero(d) <= nnz <= d.size          # stored entries may themselves be zero
        # every stored value sits where the index vector says
        seg, idx = t.sparsity["dims"][-1][2], t.sparsity["dims"][-1][3]
        flat = d.reshape(-1, t.shape[-1])
        r = int(np.searchsorted(seg, nnz // 2, side="right") - 1)
        assert flat[r, idx[nnz // 2]] == t.data.reshape(-1)[nnz // 2]
        stored += nnz
    dense_elems = sum(int(np.prod(t.shape)) for t in sparse)
    assert stored < 0.5 * dense_elems                        # the model really is pruned
    from collections import Counter
    hist = Counter(op.code for op in m.ops)
    assert hist[T.DENSIFY] == 46 and hist[T.DEPTH_TO_SPACE] == 2 and hist[T.PAD] == 43
    net = fdl.Net(os.path.join(MODELS, "face_detection_full_range_sparse.tflite"), device=-1)
    text = net.describe()
    assert text.count("DEPTH_TO_SPACE x2") == 2 and "CONV 3x3/s2 3->32 in 192x192 out 96x96 pad(t1,l1) act=relu" in text
    assert "stream=1" in text and net.num_steps < 60
    net.close()
    # unsupported sparse layouts are rejected, not mis-read
    t = sparse[0]
    bad = T.Tensor(t.index, t.name, t.shape, t.dtype, t.buffer, t.data, dict(t.sparsity, traversal_order=[3, 2, 1, 0]))
    with pytest.raises(NotImplementedError):
        T.densify(bad)


def test_roi_staging_spans_cover_every_tap(hc):
    """Zero-copy host frames: roi_fill_kernel stages, per frame row, only the span of the rotated face ROI; an eye warp runs on
    the staged copy when roi_stage_covers() admits it.  Walk every destination pixel of the face warp (and of admitted eye warps)
    through the kernels' own warp_coords() and require all four taps inside the frame to be staged (csrc/glue_math.h)."""
    import ctypes as C
    from rs_face_detection_tflite_b200._lib import CRect as fdl_rect
    hc.hc_roi_stage_check.restype = C.c_longlong
    hc.hc_roi_stage_check.argtypes = [C.POINTER(fdl_rect), C.POINTER(fdl_rect), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_longlong)]
    W, H = 1920, 1080
    rng = np.random.default_rng(11)

    def rect(xc, yc, w_px, h_px, rot):
        r = fdl_rect()
        r.x_center, r.y_center, r.width, r.height, r.rotation, r.normalized = xc / W, yc / H, w_px / W, h_px / H, rot, 1
        return r

    stats = (C.c_longlong * 2)()
    faces = eyes_checked = 0
    saved = []
    for i in range(40):
        side = float(rng.uniform(60, 900))
        rot = float(rng.uniform(-np.pi, np.pi)) if i % 4 == 0 else float(np.radians(rng.uniform(-35, 35)))
        if i % 10 == 9:
            rot = 0.0
        xc, yc = float(rng.uniform(-50, W + 50)), float(rng.uniform(-50, H + 50))      # ROIs hanging over the frame edges too
        face = rect(xc, yc, side, side, rot)
        for trim in (1, 0):
            bad = hc.hc_roi_stage_check(C.byref(face), None, W, H, 192, 64, trim, stats)
            if bad == -1:
                continue
            assert bad == 0, (i, trim, bad)
            if trim:
                faces += 1
                assert stats[0] <= stats[1]
                saved.append(stats[0] / max(stats[1], 1))
        # eye ROIs the way the pipeline makes them: squares of ~0.3 x the face side near the upper half of the face ROI, same
        # rotation +- a few degrees; some are pushed outside on purpose (those must be rejected or still be covered)
        c, s_ = np.cos(rot), np.sin(rot)
        for k in range(6):
            dx, dy = float(rng.uniform(-0.55, 0.55)) * side, float(rng.uniform(-0.55, 0.2)) * side
            es = float(rng.uniform(0.15, 0.4)) * side
            eye = rect(xc + dx * c - dy * s_, yc + dx * s_ + dy * c, es, es, rot + float(np.radians(rng.uniform(-8, 8))))
            bad = hc.hc_roi_stage_check(C.byref(face), C.byref(eye), W, H, 192, 64, 1, None)
            assert bad in (-1, -2, 0), (i, k, bad)
            eyes_checked += bad == 0
    assert faces >= 30 and eyes_checked >= 60
    assert np.mean(saved) < 0.9        # the trimming does save bytes on rotated ROIs


def test_bench_h2d_accounting_matches_the_staging_kernel(hc):
    """bench.py counts the host bytes of the zero-copy ROI staging from the face ROIs (roi_rect_bytes); the count must not be
    below what roi_fill_kernel's own span function copies, and not more than a few percent above it."""
    import ctypes as C
    import bench
    from rs_face_detection_tflite_b200._lib import CRect
    hc.hc_roi_stage_check.restype = C.c_longlong
    hc.hc_roi_stage_check.argtypes = [C.POINTER(CRect), C.POINTER(CRect), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    W, H = 1920, 1080
    rng = np.random.default_rng(3)
    stats = (C.c_longlong * 2)()

    class Roi:
        pass

    for _ in range(12):
        side, rot = float(rng.uniform(150, 800)), float(np.radians(rng.uniform(-35, 35)))
        xc, yc = float(rng.uniform(100, 1800)), float(rng.uniform(100, 1000))
        r = CRect(xc / W, yc / H, side / W, side / H, rot, 1, 0)
        assert hc.hc_roi_stage_check(C.byref(r), None, W, H, 192, 64, 1, stats) == 0
        q = Roi()
        q.x_center, q.y_center, q.width, q.height, q.rotation = xc / W, yc / H, side / W, side / H, rot
        for trim, kernel_bytes in ((True, stats[0]), (False, stats[1])):
            counted = bench.roi_rect_bytes(q, trim=trim)
            assert kernel_bytes <= counted <= 1.05 * kernel_bytes + 4096, (trim, kernel_bytes, counted)
    # rows already on the device (the letterbox gather) are not counted
    rows = frozenset(bench.letterbox_rows())
    assert len(rows) == 288
    assert bench.roi_rect_bytes(q, on_device=rows) < 0.8 * bench.roi_rect_bytes(q)


# ---- the host mirrors above the C ABI (rust_shim/, include/fdl.hpp) stay in step with include/fdl.h ------------------------
def _c_prototypes():
    """name -> number of parameters, for every FDL_API function of include/fdl.h."""
    import re
    text = open(os.path.join(ROOT, "include", "fdl.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"FDL_API\s+[^;(]*?\b(fdl_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def _c_struct_fields(name):
    import re
    text = open(os.path.join(ROOT, "include", "fdl.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?\w+[\s\*]+", "", decl)          # drop the type
        fields += [re.sub(r"\[.*?\]", "", n).strip(" *") for n in names.split(",")]
    return fields


def test_rust_shim_ffi_matches_the_c_header():
    """rust_shim/ cannot be compiled here (no cargo): at least its extern block and #[repr(C)] structs must name the functions /
    fields of include/fdl.h, in order and with the same arity."""
    import re
    src = open(os.path.join(ROOT, "rust_shim", "src", "face_detection_lite", "ffi.rs")).read()
    protos = _c_prototypes()
    assert len(protos) >= 40
    fns = re.findall(r"pub fn (fdl_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", src, flags=re.S)
    assert len(fns) >= 15
    for name, args in fns:
        assert name in protos, name + " is not declared in include/fdl.h"
        n = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n == protos[name], (name, n, protos[name])
    for name, body in re.findall(r"pub struct (fdl_\w+)\s*\{(.*?)\}", src, flags=re.S):
        if not body.strip() or "_private" in body or "_opaque" in body:
            continue                                                 # opaque handle types
        rust_fields = [f.split(":")[0].replace("pub", "").strip() for f in body.split(",") if ":" in f]
        rust_fields = [f for f in rust_fields if f]
        assert rust_fields == _c_struct_fields(name), (name, rust_fields, _c_struct_fields(name))


def test_cpp_mirror_header_compiles(tmp_path):
    """include/fdl.hpp (the C++ mirror of the reference's API) is header-only: it must at least compile against include/fdl.h."""
    import subprocess
    tu = tmp_path / "tu.cc"
    tu.write_text('#include "fdl.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(tu)])


def test_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): one JSON line on stdout -- library chatter goes to
    stderr -- with the contract's keys, the metric / unit of our own arm and a zero-copy-free e2e block."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-frames", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]


def test_warp_mismatch_rate_against_opencv_svd_is_below_1e5(hc, capsys):
    """the warp may differ from cv2 on <= 1e-5 of the pixels by one level (the f64 8x8 solve: Gaussian elimination
    in glue_math.h, DECOMP_SVD -- LAPACK here -- in the reference's OpenCV).  The kernels' own per-pixel code, compiled for the
    host, over 200 seeded rotated ROIs (7.4 M pixels) against cv2: measured, printed, bounded."""
    import math
    import synth_frames
    from oracle import glue
    from rs_face_detection_tflite_b200._lib import CRect
    hc.hc_image_to_tensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(CRect), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    r = np.random.default_rng(11)
    frames = [np.ascontiguousarray(synth_frames.load_rgb("man.jpg")), np.ascontiguousarray(synth_frames.face_frame(1))]
    bad = tot = 0
    S = 192
    for k in range(200):
        img = frames[k % 2]
        h, w = img.shape[:2]
        roi = glue.Rect(r.uniform(0.2, 0.8), r.uniform(0.2, 0.8), r.uniform(0.1, 0.9), r.uniform(0.1, 0.9), r.uniform(-math.pi, math.pi), True)
        ref = glue.image_to_tensor(img, roi, (S, S), False, (0.0, 1.0), False)
        croi = CRect(roi.x_center, roi.y_center, roi.width, roi.height, roi.rotation, 1, 0)
        t, u8, pad = np.empty((S, S, 3), np.float32), np.empty((S, S, 3), np.uint8), (C.c_double * 4)()
        assert hc.hc_image_to_tensor(img.ctypes.data, w, h, C.byref(croi), S, S, 0, 0.0, 1.0, 0, t.ctypes.data, u8.ctypes.data, pad) == 0
        diff = np.abs(u8.astype(int) - ref.u8.astype(int))
        assert diff.max() <= 1
        bad += int((diff > 0).any(axis=2).sum())
        tot += S * S
        same = diff == 0
        np.testing.assert_array_equal(t[same], ref.tensor_data[same])
    with capsys.disabled():
        print("\n[warp parity, host] %d of %d pixels differ by one level from cv2 (rate %.2e, bound 1e-5)" % (bad, tot, bad / tot))
    assert bad / tot <= 1e-5


def test_hostile_model_files_never_cross_the_abi_as_exceptions_or_crashes(fdl, tmp_path):
    """Sizes in a .tflite file are attacker-controlled: negative / zero / huge tensor dimensions and random byte damage in the
    metadata must come back as FDL_ERR_MODEL (or load), never as a C++ exception through extern "C" or a crash.  Runs in a child
    process so that a crash is a failed assertion here rather than a dead test run."""
    import struct
    import subprocess
    import sys
    data = open(os.path.join(MODELS, "face_detection_back.tflite"), "rb").read()
    pat = struct.pack("<4i", 1, 128, 128, 24)
    hits = [i for i in range(0, len(data) - 16, 4) if data[i:i + 16] == pat]
    assert hits, "no [1,128,128,24] shape vector found"
    files = []
    for n, (pos, val) in enumerate([(hits[0] + 4, -1), (hits[0] + 8, 0), (hits[0] + 12, 1 << 30), (hits[len(hits) // 2] + 4, -128), (hits[-1] + 8, -(1 << 31))]):
        b = bytearray(data)
        b[pos:pos + 4] = struct.pack("<i", val)
        f = tmp_path / ("dim%d.tflite" % n)
        f.write_bytes(bytes(b))
        files.append((str(f), True))
    r = np.random.default_rng(5)
    # random damage: the flatbuffer metadata of these files sits in the last ~15 % (the weight buffers come first)
    for n in range(60):
        b = bytearray(data)
        for _ in range(int(r.integers(1, 5))):
            b[int(r.integers(int(len(b) * 0.85), len(b)))] = int(r.integers(0, 256))
        f = tmp_path / ("fuzz%d.tflite" % n)
        f.write_bytes(bytes(b))
        files.append((str(f), False))
    script = (
        "import sys; sys.path.insert(0, %r)\n"
        "import rs_face_detection_tflite_b200 as fdl\n"
        "for path, must_fail in %r:\n"
        "    try:\n"
        "        n = fdl.Net(path, device=-1); n.describe(); n.close(); ok = True\n"
        "    except fdl.FdlError as e:\n"
        "        assert e.code in (-3, -6), (path, e.code, e.message); ok = False\n"
        "    assert not (ok and must_fail), path\n"
        "print('survived')\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), files))
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "survived" in out.stdout, out.stderr[-2000:]


def test_rust_shim_keeps_the_reference_public_surface():
    """The shim must compile user code written against the reference: every public item of the reference's hot-path modules
    (names listed here from types.rs:5-246, transform.rs:15-40, face_detection.rs:89-123/:146-205, face_landmark.rs:180-232,
    iris_landmark.rs:104-158/:268/:380, utils.rs:8) has to exist in rust_shim/ with the same name."""
    shim = os.path.join(ROOT, "rust_shim", "src", "face_detection_lite")
    read = lambda f: open(os.path.join(shim, f)).read()
    want = {
        "types.rs": ["pub struct ImageTensor", "pub struct Rect", "pub fn new(x_center: f64, y_center: f64, width: f64, height: f64, rotation: f64, normalized: bool)",
                     "pub fn size(&self) -> (f64, f64)", "pub fn scaled(&self, size: (f64, f64), normalize: bool) -> Rect", "pub fn points(&self) -> Vec<(f64, f64)>",
                     "pub struct BBox", "pub fn as_tuple", "pub fn width", "pub fn height", "pub fn empty", "pub fn normalized", "pub fn area",
                     "pub fn intersect(&self, other: &BBox) -> Option<BBox>", "pub fn scale(&self, size: (f64, f64)) -> BBox",
                     "pub fn absolute(&self, size: (i32, i32)) -> BBox", "pub struct Landmark", "pub fn new(x: f64, y: f64, z: f64)",
                     "pub struct Detection", "pub data: Array2<f32>", "pub fn new(data: Vec<f32>, score: f32)", "pub fn keypoint_count(&self) -> usize",
                     "pub fn keypoint(&self, key: usize) -> (f32, f32)", "pub fn bbox(&self) -> BBox", "pub fn scaled(&self, factor: f32) -> Detection",
                     "pub fn scaled_by_image_size(&self, image_size: (i32, i32)) -> Detection"],
        "transform.rs": ["pub enum SizeMode", "SquareLong = 1", "SquareShort = 2", "impl From<i32> for SizeMode", "pub fn to_int(self) -> i32"],
        "face_detection.rs": ["pub enum FaceIndex", "impl TryFrom<i32> for FaceIndex", "pub enum FaceDetectionModel", "FullSparse = 4",
                              "pub fn new(model_type: FaceDetectionModel, model_path: Option<String>) -> Result<FaceDetection, Error>",
                              "pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Detection>, Error>"],
        "face_landmark.rs": ["pub fn face_detection_to_roi(face_detection: Detection, image_size: (i32, i32), size_mode: Option<SizeMode>) -> Result<Rect, Error>",
                             "pub fn new(model_path: Option<String>) -> Result<FaceLandmark, Error>",
                             "pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Landmark>, Error>"],
        "iris_landmark.rs": ["pub enum IrisIndex", "pub struct IrisResults", "pub fn new(contour: Vec<Landmark>, iris: Vec<Landmark>) -> Self",
                             "pub fn eyeball_contour(&self) -> Vec<Landmark>", "pub fn new(model_path: Option<String>) -> Result<IrisLandmark, Error>",
                             "pub fn infer(&self, image: &Mat, roi: Option<Rect>, is_right_eye: Option<bool>) -> Result<IrisResults, Error>",
                             "pub fn iris_roi_from_face_landmarks(face_landmarks: Vec<Landmark>, image_size: (i32, i32)) -> Result<(Rect, Rect), Error>",
                             "pub fn update_face_landmarks_with_iris_results("],
        "utils.rs": ["pub fn convert_image_to_mat(im_bytes: &[u8]) -> Result<Mat, Error>"],
        "render.rs": ["pub struct Color", "pub fn new(r: Option<i32>, g: Option<i32>, b: Option<i32>, a: Option<i32>) -> Self", "pub struct Colors",
                      "pub const PINK: Color", "pub struct Point", "pub struct RectOrOval", "pub struct FilledRectOrOval", "pub struct Line",
                      "pub enum AnnotationData", "pub struct Annotation",
                      "pub fn new(data: Vec<AnnotationData>, normalized_positions: bool, thickness: f64, color: Color) -> Self",
                      "pub fn scaled(&self, factor: (f64, f64)) -> Result<Self, Error>", "pub fn detections_to_render_data(",
                      "pub fn landmarks_to_render_data(",
                      "pub fn render_to_image(annotations: &Vec<Annotation>, image: &DynamicImage, _blend_mode: Option<bool>) -> DynamicImage"],
    }
    for f, items in want.items():
        src = re.sub(r"\s+", " ", read(f))
        for it in items:
            assert re.sub(r"\s+", " ", it) in src, (f, it)
    lib = open(os.path.join(ROOT, "rust_shim", "src", "lib.rs")).read()
    for m in ("ffi", "types", "transform", "utils", "face_detection", "face_landmark", "iris_landmark", "render"):
        assert "pub mod %s;" % m in lib
    assert "pub fn face_landmarks_to_render_data(" in read("face_landmark.rs") and "pub const FACE_LANDMARK_CONNECTIONS" in read("face_landmark.rs")
    assert "pub fn eye_landmarks_to_render_data(" in read("iris_landmark.rs") and "pub fn iris_landmarks_to_render_data(" in read("iris_landmark.rs")


def test_tail_chain_is_planned_for_the_landmark_and_iris_graphs(fdl):
    """chain.h: the steps on maps of at most 8 x 8 pixels become one launch per chain (chain_kernel.cu).  The plan is host data: the
    FaceMesh graph chains its 6 x 6 / 3 x 3 blocks (3 faces per 128-row group); the iris graph its 8 x 8 bottlenecks (2 eyes per group)
    and, in a second chain, its 4 x 4 .. 2 x 2 ones (8 eyes per group: an op costs the same whether its rows are full or not); both
    branches of each;
    the detectors have no such tail.  Every GEMM reads a P16 operand, every depthwise an F32 one, the chains are contiguous and every
    program fits the kernel's fixed tables."""
    import re
    want = {"face_landmark": [(13, 22, 3)], "iris_landmark": [(21, 32, 2), (33, 52, 8)]}
    pat = r"chain: steps #(\d+)\.\.#(\d+) -> one launch, (\d+) items per group, (\d+) ops, (\d+) weight chunks, (\d+) parameter blocks"
    for name, chains in want.items():
        d = fdl.Net(MODELS + "/%s.tflite" % name, -1).describe()
        found = re.findall(pat, d)
        assert [(int(m[0]), int(m[1]), int(m[2])) for m in found] == chains, (name, found)
        for m in found:
            assert int(m[3]) <= 64 and int(m[4]) + int(m[5]) <= 96
            assert int(m[5]) == int(m[1]) - int(m[0]) + 1                # one parameter block per chained step
        ops = [l.split() for l in d.splitlines() if l.startswith("  ") and l.split()[0] in ("LOAD", "STORE", "POOL", "GATHER", "DW", "GEMM")]
        assert len(ops) == sum(int(m[3]) for m in found)
        for o in ops:
            if o[0] == "GEMM":
                assert "(P16" in o[3], o
            if o[0] == "DW":
                assert "(F32" in o[3], o
        assert sum(1 for o in ops if o[0] == "STORE") >= 2 * len(chains)   # both graph branches leave every chain
    for name in ("face_detection_back", "face_detection_short_range", "face_detection_full_range", "face_detection_full_range_sparse"):
        assert "chain:" not in fdl.Net(MODELS + "/%s.tflite" % name, -1).describe()
