mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log )
( timeout 300 python tools/time_variants.py 2304 1920 main main > gpurun_out/r02_variants2.log 2>&1 )
( timeout 500 python bench.py --steps 24 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err )
tail -4 gpurun_out/r02_pytest6.log | cut -c1-300; cat gpurun_out/r02_variants2.log; cut -c1-300 gpurun_out/r02_bench_c.json
