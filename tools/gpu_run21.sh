# multi-GPU tests + bench lines at N GPUs after the PARITY-build changes of the reference pin (usage: N=2 bash tools/gpu_run21.sh)
mkdir -p gpurun_out
N=${N:-2}
( timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n$N.log )
tail -4 gpurun_out/r02_pytest_n$N.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r02_bench_conus_n$N.json 2> gpurun_out/r02_bench_conus_n$N.err )
grep '^{' gpurun_out/r02_bench_conus_n$N.json | cut -c1-200
