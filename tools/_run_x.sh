mkdir -p gpurun_out
for pm in 0 64 100; do
echo "persist MB $pm"
SNP_L2_PERSIST_MB=$pm timeout 900 python tools/class_bench.py --blocks 262144 --variants 8,8c6,8c7 --small "" --classes text,mix --out gpurun_out/r02_class_bench_p$pm.json 2>&1 | grep -v "^$" | tail -3
done
