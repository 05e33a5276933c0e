mkdir -p gpurun_out
N=$1
( timeout 900 python -m pytest tests/test_multigpu.py tests/test_domain_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n$N.log )
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C5 --steps 20 --warmup 5 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err )
( NOAHMP_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_n$N.json 2> gpurun_out/r02_bench_c3_n$N.err )
tail -4 gpurun_out/r02_pytest_n$N.log | cut -c1-300; cut -c1-260 gpurun_out/r02_bench_c5_n$N.json; tail -3 gpurun_out/r02_bench_c5_n$N.err | cut -c1-300; cut -c1-260 gpurun_out/r02_bench_c3_n$N.json; grep -v trace gpurun_out/r02_bench_c3_n$N.err | tail -3 | cut -c1-300
