"""Stage timings of the pipeline with pinned host frames, copy vs zero-copy: python tools/zc_probe.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = synth_frames.face_frames(8)
host = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
for i in range(B):
    host[i] = torch.from_numpy(base[i % 8])
names = ("h2d", "det_pre", "det_net", "ssd_post", "face_warp", "lmk_net", "lmk_post_eye_warp", "iris_net", "iris_post", "d2h")
for zc in (False, True):
    for li in ((True, True), (True, False), (False, False)):
        p = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=B, run_landmarks=li[0], run_iris=li[1], model_dir="models", zero_copy_host=zc)
        for _ in range(3):
            p.collect_raw(p.submit(host))
        st = np.zeros(10)
        for _ in range(5):
            p.collect_raw(p.submit(host))
            st += np.array(p.stage_ms)
        st /= 5
        print("zero_copy" if zc else "copy     ", "lmk=%d iris=%d" % li, "total %.2f ms |" % st.sum(), " ".join("%s=%.2f" % (n, v) for n, v in zip(names, st) if v > 0.005), flush=True)
        p.close()
