"""Timeline of block_ws_kernel's first CTAs (variant build 'trace'): FDL_LIB=.../libfdl_b200_trace.so python tools/ws_trace.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_face_detection_tflite_b200 as fdl
from rs_face_detection_tflite_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
name = sys.argv[2] if len(sys.argv) > 2 else "face_detection_back"
S = {"face_detection_back": 256, "face_landmark": 192, "iris_landmark": 64}[name]
net = fdl.Net("models/%s.tflite" % name, 0)
x = np.random.default_rng(0).uniform(-1, 1, (B, S, S, 3)).astype(np.float32)
ms = net.time_steps(B, 3, x)
print("step times (us):", [round(1e3 * float(v), 1) for v in ms[:9]])
# the trace buffer now holds the LAST block_ws launch of the pass; run the net only up to step 2 is not possible, so read what is there
n = 4 * 48 * 8
buf = (ctypes.c_ulonglong * n)()
f = _lib.lib().fdl_debug_ws_trace
f.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = f(buf, n)
assert rc == 0, "not a trace build"
t = np.array(buf[:], np.int64).reshape(4, 48, 8)
names = ["load", "dw_sees", "dw0_done", "dw_all", "mma_iss", "epi_sees", "epi_done", "store"]
for cta in range(2):
    t0 = t[cta, 0, 0]
    print("CTA", cta, "(ns from its first load issue)")
    print("tile " + " ".join(n.rjust(9) for n in names))
    for it in range(24):
        print("%4d " % it + " ".join(("%9d" % (v - t0)) if v else "        -" for v in t[cta, it]))
