set -x
mkdir -p gpurun_out
# level-0 kernels of one 512^3 V-cycle, cold L2 (default cache control): first 13 launches = zero, band x3, jacobi, band x3, residual, restrict ...
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 14 -o gpurun_out/r11_sweep512_down python scripts/profile_sweep.py 512 1 > gpurun_out/r11_ncu_a.log 2>&1
# the up-stroke's prolongation at level 0 and the band sweeps after it
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_prolong" -c 7 -o gpurun_out/r11_sweep512_prolong python scripts/profile_sweep.py 512 1 > gpurun_out/r11_ncu_b.log 2>&1
