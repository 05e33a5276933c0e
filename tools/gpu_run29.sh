mkdir -p gpurun_out
N=${N:-8}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config C5 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err )
grep '^{' gpurun_out/r02_bench_c5_n$N.json | cut -c1-220
