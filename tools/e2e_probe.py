"""e2e frames/s of the zero-copy pipeline (pinned host frames, 3 batches in flight): python tools/e2e_probe.py [B] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
base = synth_frames.face_frames(8)
bufs = []
for k in range(3):
    h = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
    for i in range(B):
        h[i] = torch.from_numpy(base[(i + k) % 8])
    bufs.append(h)
p = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=B, model_dir="models", zero_copy_host=True)
for i in range(3):
    p.collect_raw(p.submit(bufs[i % 3]))
names = ("h2d", "det_pre", "det_net", "ssd_post", "face_warp", "lmk_net", "lmk_post_eye_warp", "iris_net", "iris_post", "d2h")
print("serial stage ms:", " ".join("%s=%.2f" % (n, v) for n, v in zip(names, p.stage_ms) if v > 0.005), flush=True)
for inflight in (1, 2, 3, 4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pend = []
    for s in range(steps):
        pend.append(p.submit(bufs[s % 3]))
        if len(pend) == inflight:
            p.collect_raw(pend.pop(0))
    while pend:
        p.collect_raw(pend.pop(0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("inflight %d: %.1f frames/s (%.2f ms/batch)" % (inflight, B * steps / dt, 1e3 * dt / steps), flush=True)
n_faces = sum(p._frames[i].n_faces for i in range(B))
print("faces in last batch:", n_faces)
p.close()
