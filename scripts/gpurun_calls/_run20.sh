mkdir -p gpurun_out
GMG_PRINT_STATS=1 timeout 300 python scripts/profile_step.py 256 1 > gpurun_out/r20_stats256.txt 2>&1
GMG_PRINT_STATS=1 timeout 300 python scripts/profile_sweep.py 512 1 > gpurun_out/r20_stats512.txt 2>&1
