"""frames/s through ONE process and ONE application thread over all visible GPUs (fdl_pool): python tools/pool_bench.py [B] [steps] [mode]
mode: jpeg (default; compressed frames from one pinned arena per batch slot) or zc (zero-copy pinned 1080p frames)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mode = sys.argv[3] if len(sys.argv) > 3 else "jpeg"
ngpu = fdl.device_count()
devices = [int(x) for x in os.environ.get("POOL_DEVICES", ",".join(str(i) for i in range(ngpu))).split(",")]
uniq = 16
base = synth_frames.face_frames(uniq)
pool = fdl.Pool(devices, fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=B, max_faces=1, model_dir="models", zero_copy_host=(mode == "zc"))
depth = pool.depth
if mode == "jpeg":
    files = [cv2.imencode(".jpg", np.ascontiguousarray(f[:, :, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for f in base]
    lens = [len(files[i % uniq]) for i in range(B)]
    offs = np.concatenate([[0], np.cumsum([(l + 63) & ~63 for l in lens])])
    def make():
        a = torch.zeros(int(offs[-1]), dtype=torch.uint8).pin_memory()
        for i in range(B):
            a[offs[i]:offs[i] + lens[i]] = torch.frombuffer(bytearray(files[i % uniq]), dtype=torch.uint8)
        return (a, offs[:-1], lens)
    slots = [make() for _ in range(min(depth, 8))]
    submit = pool.submit_jpeg
else:
    def make():
        h = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
        for i in range(B):
            h[i] = torch.from_numpy(base[i % uniq])
        return h
    slots = [make() for _ in range(min(depth, 8))]
    submit = pool.submit
pending = []
def pump(n):
    for s in range(n):
        pending.append(submit(slots[s % len(slots)]))
        if len(pending) == depth:
            pool.collect_raw(pending.pop(0))
    while pending:
        pool.collect_raw(pending.pop(0))
pump(2 * depth)
t0 = time.perf_counter()
pump(steps)
dt = time.perf_counter() - t0
faces = sum(pool._frames[i].n_faces for i in range(B))
print("pool devices", devices, "mode", mode, "B", B, "steps", steps, "frames/s %.0f" % (B * steps / dt), "ms/step/device %.2f" % (1e3 * dt * len(devices) / steps),
      "faces in last batch", faces)
pool.close()
