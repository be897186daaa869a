#!/bin/bash
O=gpurun_out/${1:-r01y}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}.txt 2>&1
  FDL_BLOCK_TC_MAX_N=127 timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}_maxn127.txt 2>&1
done
tail -3 $O/pytest.log; cat $O/net_bench.txt
Q="--steps 10 --warmup 3 --no-cpu-baseline --latency-iters 0 --no-zero-copy"
for sms in 148 140 132 124; do
  FDL_PERSIST_SMS=$sms timeout 300 python bench.py $Q > $O/bench_sms$sms.json 2> $O/bench_sms$sms.err
  python - "$O/bench_sms$sms.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), round(d['serial_ms_per_step'],3))
except Exception as e: print('ERR', e)
PY
done
