mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log )
( timeout 300 python tools/fast_accuracy.py gpurun_out/r02_fast_accuracy.json > gpurun_out/r02_fast_accuracy.log 2>&1 )
( timeout 420 python bench.py --steps 20 --warmup 5 --full-day > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err )
( timeout 200 python bench.py --config C5 --grid 1152 960 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_c5_small.json 2> gpurun_out/r02_bench_c5_small.err )
tail -3 gpurun_out/r02_pytest1.log; tail -2 gpurun_out/r02_fast_accuracy.log; cut -c1-600 gpurun_out/r02_bench_a.json; tail -2 gpurun_out/r02_bench_a.err; cut -c1-300 gpurun_out/r02_bench_c5_small.json; tail -3 gpurun_out/r02_bench_c5_small.err
