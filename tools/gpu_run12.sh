mkdir -p gpurun_out
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:land_kernel -s 18 -c 1 -o gpurun_out/r02_land_midday_final -f python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_midday_final.log 2>&1 )
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:land_kernel -s 6 -c 1 -o gpurun_out/r02_land_night_final -f python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_night_final.log 2>&1 )
( timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"land_kernel|glacier_kernel|seaice|permute|bin_key|DeviceRadixSort|budget|scatter_kernel|gather" --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_launches_final.log 2>&1 )
ls -la gpurun_out/*final*
