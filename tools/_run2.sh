mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_all.log; tail -5 gpurun_out/r02_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
