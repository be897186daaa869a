#!/bin/bash
O=gpurun_out/${1:-r01ai}
mkdir -p $O
for sk in 0 64 192 1088 4160 33344; do
  echo "SKEW=$sk" >> $O/out.txt
  FDL_ARENA_SKEW=$sk timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "total|#1 |#2 |#9 " >> $O/out.txt
done
cat $O/out.txt
