#!/bin/bash
O=gpurun_out/${1:-r01az}
mkdir -p $O
for m in 0 4; do
  echo "ZC_CROP=1 MARGIN=$m" >> $O/e2e.txt
  FDL_ZC_MARGIN=$m timeout 300 python tools/e2e_probe.py 256 12 2>&1 | grep -E "serial|inflight 4" >> $O/e2e.txt
done
cat $O/e2e.txt
