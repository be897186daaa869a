#!/bin/bash
O=gpurun_out/${1:-r01ae}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest_net.log 2>&1; echo "pytest exit $?" >> $O/pytest_net.log
tail -4 $O/pytest_net.log
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}.txt 2>&1
done
cat $O/net_bench.txt; head -12 $O/steps_face_detection_back.txt
