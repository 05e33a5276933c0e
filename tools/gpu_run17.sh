mkdir -p gpurun_out
( timeout 500 python tools/time_variants.py 2304 1920 main pf296 pf148 pf600 main > gpurun_out/r02_variants6_tileprefetch.log 2>&1 )
cat gpurun_out/r02_variants6_tileprefetch.log
