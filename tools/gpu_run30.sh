mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest30.log 2>&1 ); tail -n 6 gpurun_out/r02_pytest30.log
( NOAHMP_B200_TRACE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench30.json 2> gpurun_out/bench30.err ); grep '^{' gpurun_out/r02_bench30.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['step_ms'])"; grep -c "re-binning skipped" gpurun_out/bench30.err; grep "re-binning skipped" gpurun_out/bench30.err | head -3
( NOAHMP_B200_REBIN_MIN_CHANGED=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('always permute:', d['ms_per_step'])" )
