#!/bin/bash
O=gpurun_out/${1:-r01ah}
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_ws -s 2 -c 1 -o $O/prof_block_ws_128 \
    python tools/net_bench.py face_detection_back 256 1 1 > $O/ncu_block_ws.log 2>&1
tail -3 $O/ncu_block_ws.log
