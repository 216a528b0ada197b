set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_batch" > gpurun_out/r02_pytest_pack.log 2>&1; tail -3 gpurun_out/r02_pytest_pack.log
timeout 1200 python tools/compress_bench.py --blocks 65536 --variants k3c8,k3c4,k3c2,k3c1,k6c8,k6c4,k6c3,k6c2,k6c1 --out gpurun_out/r02_compress_bench1.json > gpurun_out/r02_compress_bench1.log 2>&1
cat gpurun_out/r02_compress_bench1.log
