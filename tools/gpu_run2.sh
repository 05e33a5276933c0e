mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log )
( NOAHMP_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_trace_bench.json 2> gpurun_out/r02_trace.log )
( timeout 400 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe.log 2>&1 )
( timeout 300 python tools/time_variants.py 2304 1920 main smem noesat main > gpurun_out/r02_variants1.log 2>&1 )
tail -5 gpurun_out/r02_pytest2.log; grep -c trace gpurun_out/r02_trace.log; cat gpurun_out/r02_e2e_probe.log | tail -8; cat gpurun_out/r02_variants1.log
