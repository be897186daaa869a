"""Device JPEG decode throughput: python tools/jpeg_bench.py [B] [iters] [quality]  (1080p G2 frames, decode to device memory)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
q = int(sys.argv[3]) if len(sys.argv) > 3 else 90
uniq = 16
base = synth_frames.face_frames(uniq)
files = [cv2.imencode(".jpg", np.ascontiguousarray(f[:, :, ::-1]), [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes() for f in base]
lens = [len(files[i % uniq]) for i in range(B)]
offs = np.concatenate([[0], np.cumsum([(l + 63) & ~63 for l in lens])])
arena = torch.zeros(int(offs[-1]), dtype=torch.uint8).pin_memory()
for i in range(B):
    arena[offs[i]:offs[i] + lens[i]] = torch.frombuffer(bytearray(files[i % uniq]), dtype=torch.uint8)
dec = fdl.JpegDecoder(0)
arg = (arena, offs[:-1], lens)
out = dec.decode_to_device(arg)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    out = dec.decode_to_device(arg)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
ref = cv2.cvtColor(cv2.imdecode(np.frombuffer(files[3], np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
got = out[0][out[1][3]:out[1][3] + 1080 * 1920 * 3].cpu().numpy().reshape(1080, 1920, 3)
import ctypes
from rs_face_detection_tflite_b200 import _lib
ph = (ctypes.c_longlong * 8)()
_lib.lib().fdl_debug_jpeg_phases.argtypes = [ctypes.c_void_p]
if _lib.lib().fdl_debug_jpeg_phases(ph) == 0:
    names = {0: "tables", 2: "round0", 3: "handover", 4: "scans", 5: "output"}
    print("entropy CTA0 phases (us):", {n: round((ph[i + 1] - ph[i]) / 1e3, 1) for i, n in names.items()}, "rounds", ph[7])
print("B", B, "q", q, "compressed MB/batch %.1f" % (sum(lens) / 1e6), "ms/batch (incl. host plan + H2D + sync) %.3f" % (1e3 * dt),
      "frames/s %.0f" % (B / dt), "exact", bool((ref == got).all()))
