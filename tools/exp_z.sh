#!/bin/bash
O=gpurun_out/${1:-r01z}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
for bs in 0 1; do
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  echo "BRANCH_STREAMS=$bs" >> $O/net_bench.txt
  FDL_BRANCH_STREAMS=$bs timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
done
done
Q="--steps 10 --warmup 3 --no-cpu-baseline --latency-iters 0 --no-zero-copy"
for bs in 0 1; do
  FDL_BRANCH_STREAMS=$bs timeout 300 python bench.py $Q > $O/bench_bs$bs.json 2> $O/bench_bs$bs.err
  python - "$O/bench_bs$bs.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), round(d['serial_ms_per_step'],3))
except Exception as e: print('ERR', e)
PY
done
tail -3 $O/pytest.log; cat $O/net_bench.txt
