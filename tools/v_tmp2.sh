timeout 900 python -m pytest tests/test_gpu_pool.py -x -q 2>&1 | tail -15
python tools/pool_bench.py 256 24 jpeg
POOL_DEVICES=0,0 python tools/pool_bench.py 256 24 jpeg
python tools/pool_bench.py 256 24 zc
