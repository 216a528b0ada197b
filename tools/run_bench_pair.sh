#!/bin/bash
# Runs both bench arms the way the driver does at N = 1 and keeps the lines (gpurun_out/ is scratch; copy what is cited).
mkdir -p gpurun_out
tag=${1:-final}
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_$tag.json 2> gpurun_out/r02_bench_ref_$tag.err; tail -2 gpurun_out/r02_bench_ref_$tag.err; cut -c1-400 gpurun_out/r02_bench_ref_$tag.json
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err; tail -2 gpurun_out/r02_bench_$tag.err
python - <<PY
import json
j=json.loads(open('gpurun_out/r02_bench_$tag.json').read().strip().split('\n')[-1])
r=json.loads(open('gpurun_out/r02_bench_ref_$tag.json').read().strip().split('\n')[-1])
print('decompress', j['value'], 'frac', j['roofline']['frac'], 'traffic', j['roofline']['traffic'], 'e2e', j['e2e']['value'], 'devout', j['e2e']['compressed_in_device_out']['value'], 'cpu', j['cpu_baseline']['value'], 'ref arm', r['value'])
c=j['compress']; print('compress', c['value'], 'frac', c['roofline']['frac'], 'e2e', c['e2e']['value'], 'cpu', c['cpu_baseline']['value'], 'ref arm', r['compress']['value'])
print('clocks', j['clocks'], 'launches', j['gpu_launches'])
PY
