"""The tensor-core stem against the FFMA stem: python tools/stem_probe.py MODEL BATCH"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_face_detection_tflite_b200 as fdl
SIZES = {'face_detection_back': 256, 'face_landmark': 192, 'iris_landmark': 64, 'face_detection_full_range': 192, 'face_detection_short_range': 128}
name, B = sys.argv[1], int(sys.argv[2])
S = SIZES[name]
x = np.random.default_rng(0).uniform(-1, 1, (B, S, S, 3)).astype(np.float32)
outs = {}
for mode in (0, 1):
    net = fdl.Net('models/%s.tflite' % name, 0)
    net.set_mode(mode)
    outs[mode] = net.forward(x)
for a, b in zip(outs[0], outs[1]):
    print(name, 'B', B, 'max |mode1 - mode0| = %.3g  (max |mode0| = %.3g)' % (np.abs(a - b).max(), np.abs(a).max()))
