mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = "1" ]; then
  ( timeout 900 python bench.py > gpurun_out/r02_bench_conus_n1.json 2> gpurun_out/bench31.err )
else
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N > gpurun_out/r02_bench_conus_n$N.json 2> gpurun_out/bench31_n$N.err )
fi
grep '^{' gpurun_out/r02_bench_conus_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['rebinning']['permutations_rank0'], d['roofline']['frac'], d['compute_roofline']['frac'])"
