"""Run the device-resident pipeline a few times (for ncu): python tools/pipe_once.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rs_face_detection_tflite_b200 as fdl
import synth_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
base = synth_frames.face_frames(8)
host = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8)
for i in range(B):
    host[i] = torch.from_numpy(base[i % 8])
dev = host.cuda()
p = fdl.Pipeline(fdl.FaceDetectionModel.BackCamera, (1920, 1080), max_batch=B, model_dir="models")
for _ in range(iters):
    p.collect_raw(p.submit(dev))
print("stage ms", [round(v, 3) for v in p.stage_ms])
p.close()
