#!/bin/bash
# One GPU-box visit.  Usage (under gpurun): bash tools/visit.sh <tag> [pytest -k expression] [bench: 0/1] [sanitize: 0/1]
TAG=${1:-r02}; K=${2:-}; BENCH=${3:-1}; SAN=${4:-0}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" > $O/pytest_gpu.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
fi
echo "pytest exit $?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
if [ "$BENCH" = "1" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
  timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
fi
if [ "$SAN" = "1" ]; then timeout 2400 bash tools/sanitize.sh $O/sanitize; fi
