N=$1; O=gpurun_out/r02_n$N; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
python tools/pool_bench.py 256 $((12*N)) jpeg 2>&1 | tail -1 | tee $O/pool_jpeg.txt
python tools/pool_bench.py 256 $((12*N)) zc 2>&1 | tail -1 | tee $O/pool_zc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; cut -c1-1500 $O/bench.json; tail -2 $O/bench.err
