mkdir -p gpurun_out
timeout 900 python tools/e2e_compress_trace.py 16384 > gpurun_out/r02_e2e_compress_trace.log 2>&1; tail -12 gpurun_out/r02_e2e_compress_trace.log
timeout 900 python tools/e2e_compress_probe.py 32768 > gpurun_out/r02_e2e_compress.log 2>&1; cat gpurun_out/r02_e2e_compress.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or chunk or pipeline or sizing" > gpurun_out/r02_pytest_host.log 2>&1; tail -3 gpurun_out/r02_pytest_host.log
