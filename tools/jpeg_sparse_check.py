"""Sparse colour conversion of JPEG batches (pipeline.cu: only the letterbox rows and the ROI spans are converted) against the pipeline
on fully decoded frames.  Run with FDL_JPEG_POISON=1: the lane's frame buffer is filled with 0xA5 before every decode, so a tap
outside the converted pixels changes a result.  python tools/jpeg_sparse_check.py  ->  prints "sparse ok <n compared>" or raises."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import rs_face_detection_tflite_b200 as fdl
import synth_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "models")


def encode(frame, q=90, extra=()):
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(frame[:, :, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, q] + list(extra))
    assert ok
    return enc.tobytes()


def decode(b):
    return cv2.cvtColor(cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)


def same(want, got):
    n = 0
    assert len(want) == len(got)
    for a, b in zip(want, got):
        assert [d.anchor for d in a.detections] == [d.anchor for d in b.detections]
        for da, db in zip(a.detections, b.detections):
            np.testing.assert_array_equal(da.data, db.data)
        assert len(a.faces) == len(b.faces)
        for fa, fb in zip(a.faces, b.faces):
            assert (fa.landmarks is None) == (fb.landmarks is None)
            if fa.landmarks is None:
                continue
            np.testing.assert_array_equal(fa.landmarks, fb.landmarks)
            for k in ("left_iris", "right_iris", "left_contour", "right_contour"):
                va, vb = getattr(fa, k), getattr(fb, k)
                assert (va is None) == (vb is None)
                if va is not None:
                    np.testing.assert_array_equal(va, vb)
            n += 1
    return n


def main():
    total = 0
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"      # under compute-sanitizer: the two-face batch only
    # one face per frame, faces of three sizes / rotations, 4:2:0 and 4:2:2, one file with restart markers
    frames = synth_frames.face_frames(8, start=40, faces=("man.jpg", "russ_cox_1.jpg", "russ_cox_2.jpg"))
    files = []
    for i in range(8):
        extra = []
        if i == 2:
            extra += [cv2.IMWRITE_JPEG_RST_INTERVAL, 120]
        if i == 5 and hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
            extra += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]
        files.append(encode(frames[i], 90 if i % 2 else 75, extra))
    decoded = np.stack([decode(b) for b in files])
    if not quick:
        ref = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=8, max_faces=1, model_dir=MODELS, device=0)
        want = ref.run(decoded)
        ref.close()
        pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=8, max_faces=1, model_dir=MODELS, device=0)
        for _ in range(3):       # every lane
            total += same(want, pipe.run_jpeg(files))
        pipe.close()
    # frames with two faces through the fan-out
    frames2 = [synth_frames.multi_face_frame(i) for i in range(4)]
    files2 = [encode(f) for f in frames2]
    dec2 = np.stack([decode(b) for b in files2])
    ref = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=4, max_faces=2, model_dir=MODELS, device=0)
    want2 = ref.run(dec2)
    ref.close()
    pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=4, max_faces=2, model_dir=MODELS, device=0)
    total += same(want2, pipe.run_jpeg(files2))
    pipe.close()
    if quick:
        print("sparse ok", total)
        return
    # a small frame size whose rows are not a multiple of 128 pixels wide
    small = [np.ascontiguousarray(synth_frames.face_frame(i, 1920, 1080)[200:920, 320:1600]) for i in range(3)]       # 1280 x 720
    files3 = [encode(f) for f in small]
    dec3 = np.stack([decode(b) for b in files3])
    ref = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1280, 720), max_batch=4, max_faces=1, model_dir=MODELS, device=0)
    want3 = ref.run(dec3)
    ref.close()
    pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1280, 720), max_batch=4, max_faces=1, model_dir=MODELS, device=0)
    total += same(want3, pipe.run_jpeg(files3))
    pipe.close()
    # a batch with a 4:4:4 file is off the colour fast path: whole frames are converted, as before
    if hasattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR"):
        files4 = [encode(small[0], 90, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]), files3[1]]
        dec4 = np.stack([decode(b) for b in files4])
        ref = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1280, 720), max_batch=4, max_faces=1, model_dir=MODELS, device=0)
        want4 = ref.run(dec4)
        ref.close()
        pipe = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1280, 720), max_batch=4, max_faces=1, model_dir=MODELS, device=0)
        total += same(want4, pipe.run_jpeg(files4))
        pipe.close()
    # the short-range detector (128 x 128 input: another row set) on its own
    pipe = fdl.Pipeline(fdl.FaceDetectionModel.Short, (1920, 1080), max_batch=8, max_faces=1, model_dir=MODELS, device=0, run_landmarks=False,
                        run_iris=False)
    a, b = pipe.run(decoded), pipe.run_jpeg(files)
    assert sum(len(x.detections) for x in a) > 0
    for x, y in zip(a, b):
        assert [d.anchor for d in x.detections] == [d.anchor for d in y.detections]
        for da, db in zip(x.detections, y.detections):
            np.testing.assert_array_equal(da.data, db.data)
    pipe.close()
    assert total >= 8 * 3
    print("sparse ok", total)


if __name__ == "__main__":
    main()
