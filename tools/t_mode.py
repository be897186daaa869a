"""Per-network forward time for the kernel modes (0 FFMA, 1 default, 2 serial TC BlazeBlock): python tools/t_mode.py [B]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rs_face_detection_tflite_b200 as fdl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
modes = [int(m) for m in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2, 1]
for name,S in (('face_detection_back',256),('face_landmark',192),('iris_landmark',64),('face_detection_full_range',192),('face_detection_short_range',128)):
    net = fdl.Net('models/%s.tflite'%name, 0)
    x = np.random.default_rng(0).uniform(-1,1,(B,S,S,3)).astype(np.float32)
    outs={}
    for mode in modes:
        net.set_mode(mode)
        ms = net.time_forward(B, 5, x)
        outs[mode]=net.forward(x)
        print(name, 'mode',mode,'%.3f ms/pass (B=%d)'%(ms,B), flush=True)
    if len(modes) > 1:
        print('   max abs diff', [float(np.abs(a-b).max()) for a,b in zip(outs[modes[0]],outs[modes[1]])], [float(np.abs(a).max()) for a in outs[modes[0]]])
