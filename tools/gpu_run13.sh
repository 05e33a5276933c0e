mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "library_nccl" > gpurun_out/r02_pytest_n8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n8.log )
tail -5 gpurun_out/r02_pytest_n8.log
