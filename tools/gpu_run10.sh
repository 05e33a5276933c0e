mkdir -p gpurun_out
( timeout 500 python tools/time_variants.py 2304 1920 main split2 split split4 > gpurun_out/r02_variants4_split.log 2>&1 )
cat gpurun_out/r02_variants4_split.log
