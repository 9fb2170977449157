set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r12_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
timeout 900 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r12_sweep512.json 2> gpurun_out/r12_sweep512.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r12_bench_ref.json 2> gpurun_out/r12_bench_ref.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r12_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r12_ncu_bench.log 2>&1
