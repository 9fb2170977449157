set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r33_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r33_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r33_bench.json 2> gpurun_out/r33_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r33_bench_ref.json 2> gpurun_out/r33_bench_ref.err
timeout 600 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r33_sweep512.json 2> gpurun_out/r33_sweep512.err
