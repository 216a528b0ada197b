mkdir -p gpurun_out
timeout 900 python tools/class_bench.py --blocks 65536 --variants 7 --small "" --out gpurun_out/r02_class_bench8.json > gpurun_out/r02_class_bench8.log 2>&1
cat gpurun_out/r02_class_bench8.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --decompress-only > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_b.json').read().strip().split('\n')[-1])
print(j['value'], j['roofline']['frac'], j['e2e'])
PY
