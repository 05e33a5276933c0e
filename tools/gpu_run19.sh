mkdir -p gpurun_out
( timeout 900 python bench.py --impl reference > gpurun_out/r02_reference_arm_conus.json 2> gpurun_out/ref19.err ); tail -c 1200 gpurun_out/r02_reference_arm_conus.json
( timeout 900 python bench.py > gpurun_out/r02_bench_conus_n1.json 2> gpurun_out/bench19.err ); tail -c 1500 gpurun_out/r02_bench_conus_n1.json; tail -3 gpurun_out/bench19.err
( timeout 900 python bench.py --impl reference --config C5 > gpurun_out/r02_reference_arm_c5.json 2>> gpurun_out/ref19.err ); tail -c 600 gpurun_out/r02_reference_arm_c5.json
nproc; lscpu | grep -E "Model name|Thread|Core|Socket" 
