#!/bin/bash
O=gpurun_out/pre; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i2t_rows -s 2 -c 1 -o $O/prof_i2t_rows \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-zero-copy --latency-iters 0 > $O/ncu_rows.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:i2t_kernel -s 4 -c 2 -o $O/prof_i2t_warp \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-zero-copy --latency-iters 0 > $O/ncu_warp.log 2>&1
ls -la $O
