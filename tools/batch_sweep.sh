#!/bin/bash
for b in 16 32 48 64 128 256; do timeout 100 python tools/net_bench.py face_detection_back $b 1 30; done
for b in 32 64 128 256; do timeout 100 python tools/net_bench.py face_landmark $b 1 30; done
for b in 64 128 256 512; do timeout 100 python tools/net_bench.py iris_landmark $b 1 30; done
timeout 100 python tools/step_times.py face_detection_back 32 1 5 | head -12
