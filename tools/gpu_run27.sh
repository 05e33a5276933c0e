# row groups per chunk that are binned separately (NOAHMP_B200_BIN_SUB): CONUS C3 and C5, device-resident step
mkdir -p gpurun_out
for S in 1 2 4 8; do
  for C in C3 C5; do
    NOAHMP_B200_BIN_SUB=$S timeout 400 python bench.py --config $C --steps 24 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bin_sub $S $C: %.3f ms/step' % d['ms_per_step'])"
  done
done > gpurun_out/r02_bin_sub.log 2>&1
cat gpurun_out/r02_bin_sub.log
