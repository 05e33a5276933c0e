mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest18.log 2>&1 ); tail -5 gpurun_out/r02_pytest18.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke18.log 2>&1 ); tail -2 gpurun_out/r02_smoke18.log
