mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "translated_reference or golden" > gpurun_out/r02_pytest26.log 2>&1 ); tail -n 5 gpurun_out/r02_pytest26.log
