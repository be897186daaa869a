O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_jpeg.py -x -q > $O/pytest.log 2>&1; tail -15 $O/pytest.log
python tools/jpeg_bench.py 256 5 90 > $O/jpeg_bench.log 2>&1; cat $O/jpeg_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/jpeg_launches.csv python tools/jpeg_bench.py 256 1 90 > /dev/null 2>&1
grep -E "jpeg_" $O/jpeg_launches.csv | awk -F'","' '{print substr($5,1,60), $NF}' | tail -4
