set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r10_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
timeout 900 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r10_sweep512.json 2> gpurun_out/r10_sweep512.err
