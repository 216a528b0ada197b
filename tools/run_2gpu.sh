set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_2gpu_devices.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_multi_2gpu.log; tail -5 gpurun_out/r02_pytest_multi_2gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -5 gpurun_out/r02_bench_2gpu.err; cat gpurun_out/r02_bench_2gpu.json
