#!/bin/bash
# Experiment visit: conv_tc chunk size (QC 8 vs auto), serial-kernel epilogue changes.
O=gpurun_out/${1:-r01w}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
for net in face_detection_back face_landmark iris_landmark face_detection_full_range face_detection_short_range; do
  B=256; [ $net = iris_landmark ] && B=512
  for qc in 8 0 4; do
    echo "QC=$qc" >> $O/net_bench.txt
    FDL_CONV_QC=$qc timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  done
  timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}.txt 2>&1
done
tail -3 $O/pytest.log; cat $O/net_bench.txt
