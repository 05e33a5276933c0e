mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_memcheck_resident.log python -m pytest tests/test_resident_api_gpu.py tests/test_init.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_resident.out 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_resident.log )
tail -n 4 gpurun_out/r02_sanitizer_memcheck_resident.log; tail -n 3 gpurun_out/r02_sanitizer_memcheck_resident.out
