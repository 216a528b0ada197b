mkdir -p gpurun_out
free -g | head -2
timeout 1500 python bench.py --workload frame --steps 2 --frame-gib 16 > gpurun_out/r02_bench_frame16.json 2> gpurun_out/r02_bench_frame16.err; tail -3 gpurun_out/r02_bench_frame16.err; cat gpurun_out/r02_bench_frame16.json | cut -c1-1200
