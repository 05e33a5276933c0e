# re-binning interval on the diurnal workload (quarter CONUS, 48 timed steps so that every interval re-bins at least once)
mkdir -p gpurun_out
for R in 0 6 12 20 30 48; do
  NOAHMP_B200_REBIN=$R timeout 300 python bench.py --grid 2304 1920 --steps 48 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rebin every $R steps: %.3f ms/step' % d['ms_per_step'])"
done > gpurun_out/r02_rebin_interval.log 2>&1
cat gpurun_out/r02_rebin_interval.log
