mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_cpp_driver.py tests/test_init.py tests/test_parity_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "cpp or rejects or device_loop or specialised" > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest14.log )
tail -15 gpurun_out/r02_pytest14.log | cut -c1-300
