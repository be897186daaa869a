#!/bin/bash
O=gpurun_out/${1:-r01x}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}.txt 2>&1
done
Q="--steps 10 --warmup 3 --no-cpu-baseline --latency-iters 0"
timeout 400 python bench.py $Q > $O/bench.json 2> $O/bench.err
tail -5 $O/pytest.log; cat $O/net_bench.txt; python - "$O/bench.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('copy_mode_value'), round(d['ms_per_step'],3), round(d['serial_ms_per_step'],3), d['stage_ms'])
PY
tail -3 $O/bench.err
