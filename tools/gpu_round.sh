#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag] [nets: 1 = also re-capture the network kernels (default), 0 = skip them]
TAG=${1:-r02}; NETS=${2:-1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-zero-copy --latency-iters 0 > $O/bench_under_ncu.log 2>&1
if [ "$NETS" = "1" ]; then
# the dominant kernel: block_ws_kernel on the detector's 128x128x24 stage (third block_ws launch of a detector pass)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_ws -s 2 -c 1 -o $O/prof_block_ws_128 \
    python tools/net_bench.py face_detection_back 256 1 1 > $O/ncu_block_ws.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tc -c 1 -o $O/prof_stem_tc \
    python tools/net_bench.py face_detection_back 256 1 1 > $O/ncu_stem.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o $O/prof_chain \
    python tools/net_bench.py iris_landmark 512 1 1 > $O/ncu_chain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -c 1 -o $O/prof_pw_tc \
    python tools/net_bench.py iris_landmark 512 1 1 > $O/ncu_pw_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_ws -s 1 -c 1 -o $O/prof_block_ws_iris32 \
    python tools/net_bench.py iris_landmark 512 1 1 > $O/ncu_block_ws_iris.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blaze_block_tc -s 0 -c 1 -o $O/prof_blaze_tc \
    python tools/net_bench.py face_detection_back 256 1 1 > $O/ncu_blaze.log 2>&1
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jpeg_ -s 5 -c 5 -o $O/prof_jpeg \
    python tools/jpeg_bench.py 256 1 90 > $O/ncu_jpeg.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cut -c1-3000 $O/bench.json; tail -3 $O/bench.err; cut -c1-600 $O/bench_ref.json
