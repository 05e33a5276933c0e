mkdir -p gpurun_out
( NOAHMP_B200_LIB=$PWD/noahmp_b200/libnoahmp_b200_split.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fast_parity_gpu.py tests/test_resident_api_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not full_size and not groundwater" > gpurun_out/r02_pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_split.log )
( timeout 400 python tools/time_variants.py 2304 1920 main split main split > gpurun_out/r02_variants3_split.log 2>&1 )
tail -6 gpurun_out/r02_pytest_split.log | cut -c1-300; cat gpurun_out/r02_variants3_split.log
