"""Per-op timeline of the tail-chain kernel (variant build 'trace'): FDL_LIB=.../libfdl_b200_trace.so python tools/chain_trace.py MODEL BATCH"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_face_detection_tflite_b200 as fdl
from rs_face_detection_tflite_b200 import _lib
SIZES = {'face_landmark': 192, 'iris_landmark': 64}
name, B = sys.argv[1], int(sys.argv[2])
S = SIZES[name]
net = fdl.Net('models/%s.tflite' % name, 0)
x = np.random.default_rng(0).uniform(-1, 1, (B, S, S, 3)).astype(np.float32)
ms = net.time_steps(B, 3, x)
ops = [l.split() for l in net.describe().splitlines() if l.startswith('  ') and l.split()[0] in ('LOAD', 'STORE', 'POOL', 'GATHER', 'DW', 'GEMM')]
MAXOPS = 64
n = 2 * (MAXOPS + 1) + MAXOPS * 4
buf = (ctypes.c_ulonglong * n)()
f = _lib.lib().fdl_debug_chain_trace
f.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert f(buf, n) == 0, "not a trace build"
t2 = np.array(buf[2 * (MAXOPS + 1):], np.int64).reshape(MAXOPS, 4)
t = np.array(buf[:2 * (MAXOPS + 1)], np.int64).reshape(2, MAXOPS + 1)
tot = {}
for g in range(2):
    if not t[g, 0]:
        continue
    print("group", g, "total %.1f us" % ((t[g, len(ops)] - t[g, 0]) / 1e3))
    for i, o in enumerate(ops):
        d = t[g, 1 + i] - t[g, i]
        tot[o[0]] = tot.get(o[0], 0) + d
        if g == 0:
            extra = ""
            if o[0] == "GEMM":
                extra = "  [issue %d, acc +%d, epilogue +%d, barrier +%d]" % (t2[i, 0] - t[g, i], t2[i, 1] - t2[i, 0], t2[i, 2] - t2[i, 1], t[g, 1 + i] - t2[i, 2])
            elif o[0] == "DW":
                extra = "  [op fetch %d, par wait +%d, setup +%d, loop +%d, barrier +%d]" % (t2[i, 3] - t[g, i], t2[i, 0] - t2[i, 3], t2[i, 1] - t2[i, 0], t2[i, 2] - t2[i, 1], t[g, 1 + i] - t2[i, 2])
            else:
                extra = "  [work %d, barrier +%d]" % (t2[i, 2] - t[g, i], t[g, 1 + i] - t2[i, 2])
            print("%3d %-6s %-4s %7d ns   %s%s" % (i, o[0], o[2], d, " ".join(o[3:6]), extra))
print({k: round(v / 1e3, 1) for k, v in tot.items()})
