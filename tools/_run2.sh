set -x
mkdir -p gpurun_out
timeout 900 python tools/small_block_bench.py --sizes 256,1024,4096,16384 --mib 2048 --variants 8,8c1,7 > gpurun_out/r02_small_blocks.log 2>&1
cat gpurun_out/r02_small_blocks.log
