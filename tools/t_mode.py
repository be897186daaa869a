import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rs_face_detection_tflite_b200 as fdl
for name,S in (('face_detection_back',256),('face_landmark',192),('iris_landmark',64),('face_detection_full_range',192),('face_detection_short_range',128)):
    net = fdl.Net('models/%s.tflite'%name, 0)
    x = np.random.default_rng(0).uniform(-1,1,(64,S,S,3)).astype(np.float32)
    outs={}
    for mode in (0,1):
        net.set_mode(mode)
        ms = net.time_forward(64, 5, x)
        outs[mode]=net.forward(x)
        print(name, 'mode',mode,'%.3f ms/pass (B=64)'%ms, flush=True)
    print('   max abs diff', [float(np.abs(a-b).max()) for a,b in zip(outs[0],outs[1])], [float(np.abs(a).max()) for a in outs[0]])
