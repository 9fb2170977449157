mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_band_brick -c 8 -o gpurun_out/r21_brick512 -f python scripts/profile_sweep.py 512 1 > gpurun_out/r21_ncu512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_band_brick -c 8 -o gpurun_out/r21_brick256 -f python scripts/profile_step.py 256 1 > gpurun_out/r21_ncu256.log 2>&1
