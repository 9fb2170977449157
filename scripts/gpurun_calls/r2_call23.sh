#!/bin/bash
# round 2, call 23 (1 GPU): where the Gauss-Seidel V-cycle spends its time (event classes per level, ncu --set full of the tile kernel)
mkdir -p gpurun_out
timeout 300 python scripts/profile_gs.py 256 > gpurun_out/r2c23_gs_events.txt 2>&1; cat gpurun_out/r2c23_gs_events.txt | tail -12
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_gauss' -c 6 -o /tmp/gs256 python scripts/profile_gs.py 256 > gpurun_out/r2c23_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_table.py /tmp/gs256.ncu-rep > gpurun_out/r2c23_gs_table.md 2>&1; cut -c1-250 gpurun_out/r2c23_gs_table.md
python scripts/ncu_source_top.py /tmp/gs256.ncu-rep 40 > gpurun_out/r2c23_gs_source.txt 2>&1; head -45 gpurun_out/r2c23_gs_source.txt | cut -c1-180
