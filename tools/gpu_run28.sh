mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest28.log 2>&1 ); tail -n 4 gpurun_out/r02_pytest28.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke28.log 2>&1 ); tail -n 2 gpurun_out/r02_smoke28.log
( timeout 900 python bench.py --config C5 > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/bench28_c5.err ); tail -c 300 gpurun_out/r02_bench_c5_n1.json
( timeout 900 python bench.py > gpurun_out/r02_bench_conus_n1.json 2> gpurun_out/bench28.err ); grep '^{' gpurun_out/r02_bench_conus_n1.json | cut -c1-220
