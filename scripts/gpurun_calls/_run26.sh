mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 22 -o gpurun_out/r26_sweep512_down -f python scripts/profile_sweep.py 512 1 > gpurun_out/r26_ncu512.log 2>&1
