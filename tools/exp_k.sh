#!/bin/bash
O=gpurun_out/${1:-r01al}
mkdir -p $O
for bh in 10 5 2 1; do
  echo "DBG=15 BOXH=$bh" >> $O/out.txt
  FDL_WS_NS=3 FDL_WS_DBG=15 FDL_WS_BOXH=$bh timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "#1 |#9 " >> $O/out.txt
done
cat $O/out.txt
