"""Serial stage times of the device-resident pipeline: python tools/stage_probe.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = synth_frames.face_frames(8)
dev = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8)
for i in range(B):
    dev[i] = torch.from_numpy(base[i % 8])
dev = dev.cuda()
names = ("h2d", "det_pre", "det_net", "ssd_post", "face_warp", "lmk_net", "lmk_post_eye_warp", "iris_net", "iris_post", "d2h")
p = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=B, model_dir="models")
for _ in range(3):
    p.collect_raw(p.submit(dev))
st = np.zeros(10)
for _ in range(5):
    p.collect_raw(p.submit(dev))
    st += np.array(p.stage_ms)
st /= 5
print("total %.3f ms |" % st[1:9].sum(), " ".join("%s=%.3f" % (n, v) for n, v in zip(names, st) if v > 0.005), flush=True)
p.close()
