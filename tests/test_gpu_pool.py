"""fdl_pool (every GPU of a box behind one handle) and fdl_frame (a frame staged once for the per-frame API), through the C ABI."""
import os

import numpy as np
import pytest

from conftest import MODELS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(a, b):
    assert [d.anchor for d in a.detections] == [d.anchor for d in b.detections]
    for da, db in zip(a.detections, b.detections):
        np.testing.assert_array_equal(da.data, db.data)
    assert len(a.faces) == len(b.faces)
    for fa, fb in zip(a.faces, b.faces):
        assert (fa.landmarks is None) == (fb.landmarks is None)
        if fa.landmarks is not None:
            np.testing.assert_array_equal(fa.landmarks, fb.landmarks)
            np.testing.assert_array_equal(fa.left_iris, fb.left_iris)
            np.testing.assert_array_equal(fa.right_contour, fb.right_contour)


def test_pool_equals_pipeline_and_balances(fdl, gpu):
    """Two pipelines (on this one GPU when the box has no second), eight tickets in flight from one thread -- raw frames and JPEG
    files mixed -- give exactly the plain pipeline's results, and the dispatcher spreads them over both workers."""
    import cv2
    import synth_frames
    n_dev = fdl.device_count()
    devices = [gpu, 1 if n_dev > 1 else gpu]
    frames = synth_frames.face_frames(12, start=60)
    files = [cv2.imencode(".jpg", np.ascontiguousarray(f[:, :, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for f in frames]
    decoded = np.stack([cv2.cvtColor(cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB) for b in files])
    ref_pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=4, max_faces=1, model_dir=MODELS, device=gpu)
    want_raw = [ref_pipe.run(frames[i:i + 3]) for i in range(0, 12, 3)]
    want_jpg = [ref_pipe.run(decoded[i:i + 3]) for i in range(0, 12, 3)]
    ref_pipe.close()
    pool = fdl.Pool(devices, fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=4, max_faces=1, model_dir=MODELS)
    assert pool.depth == 8
    tickets = []
    for k in range(4):
        tickets.append(("raw", k, pool.submit(frames[3 * k:3 * k + 3])))
        tickets.append(("jpg", k, pool.submit_jpeg(files[3 * k:3 * k + 3])))
    with pytest.raises(fdl.FdlError):
        pool.submit(frames[:1])                      # ninth ticket: every pipeline is full
    used = set()
    for kind, k, t in reversed(tickets):             # collected out of order
        got = pool.collect(t)
        used.add(pool.last_device_index)
        want = (want_raw if kind == "raw" else want_jpg)[k]
        assert len(got) == 3
        for a, b in zip(want, got):
            _same(a, b)
    assert used == {0, 1}
    with pytest.raises(fdl.FdlError):
        pool.collect(tickets[0][2])                  # a ticket is collected once
    with pytest.raises(fdl.FdlError):
        pool.run_jpeg([files[0][:len(files[0]) // 2]])   # the worker's error comes back through collect
    assert len(pool.run(frames[:2])) == 2
    pool.close()
    with pytest.raises(fdl.FdlError):
        fdl.Pool([gpu, 99], model_dir=MODELS)


def test_frame_is_uploaded_once_for_the_reference_call_sequence(fdl, gpu, man):
    """lib.rs:20-40 with an fdl_frame: detector, landmark and both iris calls read the one device copy and return exactly what they
    return for the host image; Frame(jpeg=...) is convert_image_to_mat on the device."""
    det = fdl.FaceDetection(fdl.FaceDetectionModel.BackCamera, MODELS, device=gpu)
    lmk = fdl.FaceLandmark(MODELS + "/face_landmark.tflite", device=gpu)
    iris = fdl.IrisLandmark(MODELS + "/iris_landmark.tflite", device=gpu)
    h, w = man.shape[:2]

    def chain(img):
        faces = det.infer(img)
        roi = fdl.face_detection_to_roi(faces[0], (w, h))
        lm = lmk.infer(img, roi)
        lroi, rroi = fdl.iris_roi_from_face_landmarks(lm, (w, h))
        return faces, lm, iris.infer(img, rroi, True), iris.infer(img, lroi, False)

    want = chain(man)
    jpeg = open(os.path.join(ROOT, "test_data", "man.jpg"), "rb").read()
    for frame in (fdl.Frame(man, device=gpu), fdl.Frame(jpeg=jpeg, device=gpu)):
        assert frame.size == (w, h)
        got = chain(frame)
        np.testing.assert_array_equal(got[0][0].data, want[0][0].data)
        assert [(l.x, l.y, l.z) for l in got[1]] == [(l.x, l.y, l.z) for l in want[1]]
        for a, b in ((got[2], want[2]), (got[3], want[3])):
            assert [(l.x, l.y) for l in a.contour] == [(l.x, l.y) for l in b.contour]
            assert [(l.x, l.y) for l in a.iris] == [(l.x, l.y) for l in b.iris]
        frame.close()
    f = fdl.Frame(device=gpu)
    with pytest.raises(fdl.FdlError):
        det.infer(f)                                  # nothing uploaded yet
    f.upload(man[:100, :80])
    assert f.size == (80, 100)
    with pytest.raises(fdl.FdlError):
        f.upload_jpeg(jpeg[:200])
    f.close()
    for o in (det, lmk, iris):
        o.close()
