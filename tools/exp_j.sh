#!/bin/bash
O=gpurun_out/${1:-r01ak}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest_net.log 2>&1; echo "pytest exit $?" >> $O/pytest_net.log
tail -6 $O/pytest_net.log
for cfg in "0 0" "1 0" "1 3"; do
  set -- $cfg
  echo "ROWS=$1 NS=$2" >> $O/out.txt
  FDL_WS_ROWS=$1 FDL_WS_NS=$2 FDL_WS_VERBOSE=1 timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "total|#1 |#9 |128x128 G" | sort -u >> $O/out.txt
  FDL_WS_ROWS=$1 FDL_WS_NS=$2 timeout 120 python tools/step_times.py face_landmark 256 1 10 2>&1 | grep -E "total|#1 " >> $O/out.txt
done
for d in 15 7; do
  echo "ROWS=1 NS=3 DBG=$d" >> $O/out.txt
  FDL_WS_NS=3 FDL_WS_DBG=$d timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "#1 " >> $O/out.txt
done
cat $O/out.txt
