mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --profile-from-start off -c 8 -o gpurun_out/r35_pcg256 -f python scripts/profile_step.py 256 1 > gpurun_out/r35_ncu256.log 2>&1
