mkdir -p gpurun_out
timeout 900 python tools/e2e_compress_probe.py 65536 config3 > gpurun_out/r02_e2e_compress.log 2>&1; cat gpurun_out/r02_e2e_compress.log
timeout 900 python tools/e2e_compress_probe.py 65536 mix > gpurun_out/r02_e2e_compress_mix.log 2>&1; cat gpurun_out/r02_e2e_compress_mix.log
timeout 900 python bench.py --workload frame --steps 3 --frame-gib 4 > gpurun_out/r02_bench_frame.json 2> gpurun_out/r02_bench_frame.err; tail -3 gpurun_out/r02_bench_frame.err; cat gpurun_out/r02_bench_frame.json | cut -c1-200; grep -o '"frame_compress_GBps[^,]*,[^,]*' gpurun_out/r02_bench_frame.json
