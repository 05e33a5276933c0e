# the host-buffer call at 4 GPUs: final defaults against the settings of the earlier 6.0 ms line, same box, back to back
mkdir -p gpurun_out
N=4
for tag in final old final2; do
  if [ "$tag" = "old" ]; then export NOAHMP_B200_REBIN_MIN_CHANGED=0 NOAHMP_B200_BIN_SUB=1; else unset NOAHMP_B200_REBIN_MIN_CHANGED NOAHMP_B200_BIN_SUB; fi
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r02_e2e_n4_$tag.json 2> gpurun_out/r02_e2e_n4_$tag.err )
  grep '^{' gpurun_out/r02_e2e_n4_$tag.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],2), d['e2e']['call_ms_rank0'][:6])"
done
