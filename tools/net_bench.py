"""Per-network microbenchmark (device-resident): python tools/net_bench.py MODEL BATCH MODE ITERS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_face_detection_tflite_b200 as fdl
SIZES = {'face_detection_back': 256, 'face_landmark': 192, 'iris_landmark': 64, 'face_detection_full_range': 192, 'face_detection_short_range': 128, 'face_detection_full_range_sparse': 192}
name, batch, mode, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
S = SIZES[name]
net = fdl.Net('models/%s.tflite' % name, 0)
net.set_mode(mode)
x = np.random.default_rng(0).uniform(-1, 1, (batch, S, S, 3)).astype(np.float32)
ms = net.time_forward(batch, iters, x)
print(name, 'B', batch, 'mode', mode, '%.3f ms/pass' % ms, '%.2f us/item' % (1e3 * ms / batch))
