set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r34_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r34_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 112 -o gpurun_out/r34_pcg256 -f python scripts/profile_step.py 256 1 > gpurun_out/r34_ncu256.log 2>&1
