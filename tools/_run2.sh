mkdir -p gpurun_out
timeout 900 python tools/class_bench.py --blocks 65536 --variants 7 --small "" --classes text,markup,kppkn,records,mix --out gpurun_out/r02_class_bench9.json > gpurun_out/r02_class_bench9.log 2>&1
cat gpurun_out/r02_class_bench9.log
