#!/bin/bash
O=gpurun_out/${1:-r01aq}
mkdir -p $O
for rep in 1 2; do
for lib in old new; do
  echo "LIB=$lib" >> $O/out.txt
  L=""; [ $lib = old ] && L=$PWD/rs_face_detection_tflite_b200/libfdl_old.so
  FDL_LIB=$L timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "total|#1 |#9 |#17 " >> $O/out.txt
done
done
echo "new, F16=1" >> $O/out.txt
FDL_WS_F16=1 timeout 120 python tools/step_times.py face_detection_back 256 1 10 2>&1 | grep -E "total|#1 |#9 |#17 " >> $O/out.txt
cat $O/out.txt
