set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r7_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
GMG_COARSE_FUSED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_nofused.json 2> gpurun_out/r7_bench_nofused.err
