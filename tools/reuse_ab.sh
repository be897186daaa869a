#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for v in 1 0 1 0; do echo "== FDL_ZC_TRIM=$v"; FDL_ZC_TRIM=$v timeout 200 python tools/e2e_probe.py 256 12 2>&1 | grep -v "^faces"; done
