set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
ncu --set full --clock-control none --import-source on --profile-from-start off -c 16 -o gpurun_out/r1_full_a python scripts/profile_step.py 256 1 > gpurun_out/r1_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_prolong|k_vec" -s 7 -c 8 -o gpurun_out/r1_full_b python scripts/profile_step.py 256 1 > gpurun_out/r1_ncu_b.log 2>&1
ls -la gpurun_out
