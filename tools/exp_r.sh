#!/bin/bash
O=gpurun_out/${1:-r01bc}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest_pipe.log 2>&1; echo "pytest exit $?" >> $O/pytest_pipe.log
tail -5 $O/pytest_pipe.log
for c in 0 1; do
  echo "ZC_DMA=$c" >> $O/e2e.txt
  FDL_ZC_DMA=$c timeout 300 python tools/e2e_probe.py 256 16 >> $O/e2e.txt 2>&1
done
cat $O/e2e.txt
