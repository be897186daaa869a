"""Per-launch times of one network inside whole forward passes: python tools/step_times.py MODEL BATCH [MODE] [ITERS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_face_detection_tflite_b200 as fdl
SIZES = {'face_detection_back': 256, 'face_landmark': 192, 'iris_landmark': 64, 'face_detection_full_range': 192, 'face_detection_short_range': 128, 'face_detection_full_range_sparse': 192}
name, batch = sys.argv[1], int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
S = SIZES[name]
net = fdl.Net('models/%s.tflite' % name, 0)
net.set_mode(mode)
x = np.random.default_rng(0).uniform(-1, 1, (batch, S, S, 3)).astype(np.float32)
ms = net.time_steps(batch, iters, x)
steps = [l for l in net.describe().splitlines() if l.startswith('#')]
print(name, 'B', batch, 'mode', mode, 'total %.3f ms' % ms.sum())
for t, s in zip(ms, steps):
    print('%8.1f us %5.1f%%  %s' % (1e3 * t, 100 * t / ms.sum(), s[:120]))
