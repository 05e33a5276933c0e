mkdir -p gpurun_out
( timeout 600 python tools/time_variants.py 2304 1920 main g128 g64 b512g256 main > gpurun_out/r02_variants7_barrier_groups.log 2>&1 )
cat gpurun_out/r02_variants7_barrier_groups.log
