#!/bin/bash
O=gpurun_out/preab; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for v in "FDL_I2T_PX=1" "FDL_I2T_PX=2" "FDL_I2T_PX=4"; do
  echo "== $v"; env $v timeout 200 python tools/stage_probe.py 256 2>&1 | tail -1
done
