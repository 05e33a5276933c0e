mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log )
( timeout 500 python bench.py --steps 20 --warmup 5 --full-day > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err )
tail -15 gpurun_out/r02_pytest4.log | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
