mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest11.log )
( timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke.log 2>&1 )
tail -6 gpurun_out/r02_pytest11.log | cut -c1-300; tail -3 gpurun_out/r02_smoke.log
