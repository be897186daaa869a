#!/bin/bash
O=gpurun_out/${1:-r01ac}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest_net.log 2>&1; echo "pytest exit $?" >> $O/pytest_net.log
tail -15 $O/pytest_net.log
for f in 0 1; do
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  echo "WS_F16=$f" >> $O/net_bench.txt
  FDL_WS_F16=$f timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
done
done
cat $O/net_bench.txt
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  FDL_WS_VERBOSE=1 timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}.txt 2> $O/steps_${net}.err
done
head -12 $O/steps_face_detection_back.txt; sort -u $O/steps_*.err | head -20
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
