# compute-sanitizer passes over the small end-to-end case of __graft_entry__.smoke() and the groundwater device loop
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log )
tail -5 gpurun_out/r02_sanitizer_memcheck.log
( timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_initcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_initcheck.out 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02_sanitizer_initcheck.log )
tail -12 gpurun_out/r02_sanitizer_initcheck.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_memcheck_gw.log python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "groundwater_device_loop or c1_bitexact" > gpurun_out/r02_sanitizer_memcheck_gw.out 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_gw.log )
tail -4 gpurun_out/r02_sanitizer_memcheck_gw.log; tail -3 gpurun_out/r02_sanitizer_memcheck_gw.out
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/r02_sanitizer_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log )
tail -4 gpurun_out/r02_sanitizer_racecheck.log
