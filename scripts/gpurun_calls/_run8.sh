set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r8_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
GMG_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_nopdl.json 2> gpurun_out/r8_bench_nopdl.err
# warm-cache ncu of the level-0 kernels of one PCG iteration (cache-control none: L2 stays warm as in the real pipeline)
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on --profile-from-start off -c 40 -o gpurun_out/r8_warm python scripts/profile_step.py 256 1 > gpurun_out/r8_ncu.log 2>&1
ls -la gpurun_out | tail -5
