set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r6_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
GMG_COARSE_FUSED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_nofused.json 2> gpurun_out/r6_bench_nofused.err
GMG_PRINT_STATS=1 GMG_NO_DIRECT_DMA=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_staged.json 2> gpurun_out/r6_bench_staged.err
