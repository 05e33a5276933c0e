mkdir -p gpurun_out
N=8
( NOAHMP_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_n$N.json 2> gpurun_out/r02_bench_c3_n$N.err )
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C5 --steps 20 --warmup 5 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err )
( timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_pytest_n$N.log 2>&1 )
( nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_topo_n8.txt; numactl -H >> gpurun_out/r02_topo_n8.txt 2>&1 )
grep '^{' gpurun_out/r02_bench_c3_n$N.json | cut -c1-260; grep -v trace gpurun_out/r02_bench_c3_n$N.err | tail -3 | cut -c1-300; grep '^{' gpurun_out/r02_bench_c5_n$N.json | cut -c1-260; tail -3 gpurun_out/r02_pytest_n$N.log
