set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "handmade or golden_chunks or bad_data" > gpurun_out/r02_memcheck2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck2.log
tail -4 gpurun_out/r02_memcheck2.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -5 gpurun_out/r02_pytest2.log
timeout 900 python tools/class_bench.py --blocks 65536 --variants 7,7w2048,7w2048c6,7w8192,5 --small "" --out gpurun_out/r02_class_bench2.json > gpurun_out/r02_class_bench2.log 2>&1
cat gpurun_out/r02_class_bench2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decompress_v7 -s 2 -c 1 -o gpurun_out/r02_v7b_text python tools/profile_class.py text 32768 1 > gpurun_out/r02_prof_text.log 2>&1
tail -3 gpurun_out/r02_prof_text.log
