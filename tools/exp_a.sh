#!/bin/bash
O=gpurun_out/${1:-r01aa}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -30 $O/pytest.log
timeout 120 python tools/step_times.py face_detection_full_range_sparse 256 1 5 > $O/steps_sparse.txt 2>&1; head -3 $O/steps_sparse.txt
timeout 120 python tools/step_times.py face_detection_full_range 256 1 5 > $O/steps_full.txt 2>&1; head -3 $O/steps_full.txt
