#!/bin/bash
O=gpurun_out/${1:-r01an}
mkdir -p $O
for tw in 0 32; do
  echo "STEM_TW=$tw" >> $O/out.txt
  FDL_STEM_TW=$tw timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "#0 " >> $O/out.txt
  FDL_STEM_TW=$tw timeout 120 python tools/step_times.py face_landmark 256 1 10 2>&1 | grep -E "#0 " >> $O/out.txt
  FDL_STEM_TW=$tw timeout 120 python tools/step_times.py iris_landmark 512 1 10 2>&1 | grep -E "#0 " >> $O/out.txt
  FDL_STEM_TW=$tw timeout 120 python tools/step_times.py face_detection_full_range_sparse 256 1 10 2>&1 | grep -E "total|#0 " >> $O/out.txt
done
cat $O/out.txt
