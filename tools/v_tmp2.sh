timeout 900 python -m pytest tests -m gpu -x -q -k "net or pipeline or golden" 2>&1 | tail -3
for m in face_detection_back face_landmark iris_landmark face_detection_full_range face_detection_short_range; do python tools/net_bench.py $m 256 1 10; done
python tools/net_bench.py iris_landmark 512 1 10
