set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r16_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r16_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r16_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
timeout 600 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r16_sweep512.json 2> gpurun_out/r16_sweep512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r16_ncu_bench.log 2>&1
