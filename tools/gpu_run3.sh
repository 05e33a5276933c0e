mkdir -p gpurun_out
( NOAHMP_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 15 --no-cpu-baseline > gpurun_out/r02_trace_day_bench.json 2> gpurun_out/r02_trace_day.log )
( timeout 500 python tools/e2e_probe.py > gpurun_out/r02_e2e_probe.log 2>&1 )
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:land_kernel -s 18 -c 1 -o gpurun_out/r02_land_midday -f python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_midday.log 2>&1 )
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:land_kernel -s 6 -c 1 -o gpurun_out/r02_land_night -f python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_night.log 2>&1 )
( timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"land_kernel|glacier_kernel|seaice|permute|bin_key|DeviceRadixSort|budget|scatter_kernel|gather" --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1 )
grep "trace" gpurun_out/r02_trace_day.log | tail -4; tail -8 gpurun_out/r02_e2e_probe.log; ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_midday.log; wc -l gpurun_out/r02_launches.csv
