mkdir -p gpurun_out
GMG_PRINT_STATS=1 timeout 300 python scripts/profile_step.py 256 1 2>&1 | grep "level" > gpurun_out/r22_stats256.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err
timeout 600 python bench.py --workload vcycle --size 512 --steps 5 --warmup 3 > gpurun_out/r22_sweep512.json 2> gpurun_out/r22_sweep512.err
