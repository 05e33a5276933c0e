mkdir -p gpurun_out
( timeout 500 python bench.py --steps 20 --warmup 5 --full-day > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err )
( timeout 500 python bench.py --config C5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c5_n1b.json 2> gpurun_out/r02_bench_c5_n1b.err )
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref_c3.json 2> gpurun_out/r02_ref_c3.err )
cut -c1-200 gpurun_out/r02_bench_final.json; cut -c1-200 gpurun_out/r02_bench_c5_n1b.json; cut -c1-200 gpurun_out/r02_ref_c3.json
