set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r18_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err
timeout 600 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r18_sweep512.json 2> gpurun_out/r18_sweep512.err
