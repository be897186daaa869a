#!/bin/bash
O=gpurun_out/${1:-r01ad}
mkdir -p $O
for pad in 1 0; do
  echo "PAD6=$pad" >> $O/out.txt
  FDL_WS_PAD6=$pad FDL_WS_VERBOSE=1 timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "128x128 G=|#1 |#9 |total" | sort -u >> $O/out.txt
  FDL_WS_PAD6=$pad timeout 120 python tools/net_bench.py face_detection_back 256 1 20 >> $O/out.txt 2>&1
done
cat $O/out.txt
