set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_fused8.json 2> gpurun_out/r2_bench_fused8.err
GMG_CLUSTER=16 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_fused16.json 2> gpurun_out/r2_bench_fused16.err
GMG_COARSE_FUSED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nofused.json 2> gpurun_out/r2_bench_nofused.err
GMG_FUSED_CELLS=1000000 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_fused_l1.json 2> gpurun_out/r2_bench_fused_l1.err
ls -la gpurun_out
