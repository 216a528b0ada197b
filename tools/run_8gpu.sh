mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)\|model name" >> gpurun_out/r02_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/pcie_probe_multi.py > gpurun_out/r02_pcie_8gpu.log 2>&1; grep "^{" gpurun_out/r02_pcie_8gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -3 gpurun_out/r02_bench_8gpu.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_8gpu.json').read().strip().split('\n')[-1])
print({k:j[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e',j['e2e']['value'],'numa',j.get('numa'))
print('compress',j['compress']['value'],'e2e',j['compress']['e2e']['value'])
print('roundtrip',j['roundtrip']['value'], j['nccl_scatter_gather'])
PY
SNP_BENCH_NUMA=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 3 --warmup 3 --decompress-only --blocks 262144 > gpurun_out/r02_bench_8gpu_nonuma.json 2> gpurun_out/r02_bench_8gpu_nonuma.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_8gpu_nonuma.json').read().strip().split('\n')[-1])
print('no-numa e2e',j['e2e']['value'])
PY
