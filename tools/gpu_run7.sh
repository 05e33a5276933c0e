mkdir -p gpurun_out
( timeout 500 python bench.py --config C5 --steps 20 --warmup 5 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err )
( timeout 500 python bench.py --config C4 --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err )
( timeout 300 python bench.py --impl reference --config C5 --steps 10 --warmup 2 > gpurun_out/r02_ref_c5.json 2> gpurun_out/r02_ref_c5.err )
cut -c1-300 gpurun_out/r02_bench_c5_n1.json; tail -2 gpurun_out/r02_bench_c5_n1.err; cut -c1-300 gpurun_out/r02_bench_c4_n1.json; tail -2 gpurun_out/r02_bench_c4_n1.err; cut -c1-200 gpurun_out/r02_ref_c5.json
