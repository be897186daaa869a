#!/bin/bash
# compute-sanitizer record: memcheck + racecheck + synccheck over the network kernels (mbarrier rings, TMA, tcgen05)
# and a whole small pipeline batch.  Usage (under gpurun): bash tools/sanitize.sh <outdir>
O=${1:-gpurun_out/sanitize}
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool name cmd...
  local tool=$1 name=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 "$@" > $O/${tool}_${name}.log 2>&1
  echo "$tool $name exit $? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${tool}_${name}.log | tail -1)" >> $O/summary.txt
}
for net in face_detection_back face_landmark iris_landmark face_detection_full_range; do
  run memcheck  net_$net python tools/net_bench.py $net 3 1 1
  run racecheck net_$net python tools/net_bench.py $net 3 1 1
done
run synccheck net_back python tools/net_bench.py face_detection_back 3 1 1
run synccheck net_iris python tools/net_bench.py iris_landmark 5 1 1
run memcheck  jpeg python tools/jpeg_bench.py 4 1 90
run racecheck jpeg python tools/jpeg_bench.py 4 1 90
FDL_JPEG_POISON=1 run memcheck  jpeg_sparse python tools/jpeg_sparse_check.py quick
FDL_JPEG_POISON=1 run racecheck jpeg_sparse python tools/jpeg_sparse_check.py quick
run memcheck  pipeline python tools/pipe_once.py 3 1
run racecheck pipeline python tools/pipe_once.py 3 1
cat $O/summary.txt
