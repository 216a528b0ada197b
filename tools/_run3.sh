set -x
mkdir -p gpurun_out
SNP_DECOMP_KERNEL=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decompress_v8 -s 2 -c 1 -o gpurun_out/r02_v8a_text python tools/profile_class.py text 131072 1 > gpurun_out/r02_prof_v8a.log 2>&1
tail -3 gpurun_out/r02_prof_v8a.log
