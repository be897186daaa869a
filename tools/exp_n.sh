#!/bin/bash
O=gpurun_out/${1:-r01ao}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest_net.log 2>&1; echo "pytest exit $?" >> $O/pytest_net.log
tail -4 $O/pytest_net.log
timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "total|#0 " >> $O/out.txt
timeout 120 python tools/step_times.py face_landmark 256 1 10 2>&1 | grep -E "total|#0 " >> $O/out.txt
timeout 120 python tools/step_times.py iris_landmark 512 1 10 2>&1 | grep -E "total|#0 " >> $O/out.txt
cat $O/out.txt
