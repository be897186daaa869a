#!/bin/bash
O=gpurun_out/${1:-r01aj}
mkdir -p $O
for d in 0 8 15 7; do
  echo "DBG=$d" >> $O/out.txt
  FDL_WS_DBG=$d timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "#1 |#9 " >> $O/out.txt
done
cat $O/out.txt
