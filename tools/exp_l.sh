#!/bin/bash
O=gpurun_out/${1:-r01am}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest_net.log 2>&1; echo "pytest exit $?" >> $O/pytest_net.log
tail -6 $O/pytest_net.log
for f in 0 1; do
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  echo "TC_F16=$f" >> $O/net_bench.txt
  FDL_TC_F16=$f timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  FDL_TC_F16=$f timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}_$f.txt 2>&1
done
done
cat $O/net_bench.txt
