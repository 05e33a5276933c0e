mkdir -p gpurun_out
( timeout 500 python tools/time_variants.py 2304 1920 main expopt vec r144 main > gpurun_out/r02_variants5_flags.log 2>&1 )
cat gpurun_out/r02_variants5_flags.log
