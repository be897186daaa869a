#!/bin/bash
# Experiment visit: PDL on/off, 192- vs 384-thread serial BlazeBlock kernel, batches in flight, batch size.
O=gpurun_out/${1:-r01v}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
for net in face_detection_back face_landmark iris_landmark; do
  B=256; [ $net = iris_landmark ] && B=512
  for pdl in 0 1; do for thr in 192 384; do
    echo "PDL=$pdl THR=$thr" >> $O/net_bench.txt
    FDL_PDL=$pdl FDL_TC_THREADS=$thr timeout 120 python tools/net_bench.py $net $B 1 20 >> $O/net_bench.txt 2>&1
  done; done
  FDL_TC_THREADS=384 timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}_384.txt 2>&1
  FDL_TC_THREADS=192 timeout 120 python tools/step_times.py $net $B 1 10 > $O/steps_${net}_192.txt 2>&1
done
Q="--steps 10 --warmup 3 --no-cpu-baseline --latency-iters 0 --no-zero-copy"
FDL_PDL=0 FDL_TC_THREADS=192 timeout 300 python bench.py $Q --dev-inflight 1 > $O/bench_base.json 2> $O/bench_base.err
FDL_PDL=1 timeout 300 python bench.py $Q --dev-inflight 1 > $O/bench_pdl_if1.json 2> $O/bench_pdl_if1.err
FDL_PDL=1 timeout 300 python bench.py $Q --dev-inflight 3 > $O/bench_pdl_if3.json 2> $O/bench_pdl_if3.err
FDL_PDL=0 timeout 300 python bench.py $Q --dev-inflight 3 > $O/bench_nopdl_if3.json 2> $O/bench_nopdl_if3.err
FDL_PDL=1 timeout 300 python bench.py $Q --dev-inflight 3 --batch 512 > $O/bench_pdl_if3_b512.json 2> $O/bench_pdl_if3_b512.err
tail -3 $O/pytest.log; cat $O/net_bench.txt
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['serial_ms_per_step'],3), d['stage_ms'])
except Exception as e: print('ERR', e)
PY
done
