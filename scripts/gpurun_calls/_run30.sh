mkdir -p gpurun_out
GMG_PRINT_STATS=1 timeout 300 python scripts/setup_probe2.py 256 > gpurun_out/r30_setup.txt 2>&1
