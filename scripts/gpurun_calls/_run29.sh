mkdir -p gpurun_out
GMG_PRINT_STATS=1 timeout 300 python scripts/profile_step.py 256 1 > gpurun_out/r29_stats256.txt 2>&1
nproc >> gpurun_out/r29_stats256.txt
