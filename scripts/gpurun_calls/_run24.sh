mkdir -p gpurun_out
GMG_PRINT_STATS=1 timeout 300 python scripts/profile_step.py 256 1 > gpurun_out/r24_stats256.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err
