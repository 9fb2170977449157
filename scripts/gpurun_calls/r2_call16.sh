#!/bin/bash
# round 2, call 16 (1 GPU): ncu --set full of the ring-halo tile kernel (128^3: both variants; 256^3: zero-grid variant) beside the sweep-per-launch kernels
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_band' -s 100 -c 16 -o /tmp/tiles128 python bench.py --quick --size 128 --steps 1 --warmup 3 > gpurun_out/r2c16_ncu128.log 2>&1; echo "ncu 128 rc=$?"
GMG_BAND_TILES=0 timeout 600 ncu --set full --clock-control none -k regex:'k_band' -s 240 -c 12 -o /tmp/sweeps128 python bench.py --quick --size 128 --steps 1 --warmup 3 > gpurun_out/r2c16_ncu128s.log 2>&1; echo "ncu 128 sweeps rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'k_band' -s 120 -c 20 -o /tmp/tiles256 python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2c16_ncu256.log 2>&1; echo "ncu 256 rc=$?"
python scripts/ncu_table.py /tmp/tiles128.ncu-rep /tmp/sweeps128.ncu-rep /tmp/tiles256.ncu-rep > gpurun_out/r2c16_tables.md 2>&1
python scripts/ncu_source_top.py /tmp/tiles128.ncu-rep 60 > gpurun_out/r2c16_source_top.txt 2>&1
ncu -i /tmp/tiles128.ncu-rep --page details --csv > gpurun_out/r2c16_details128.csv 2>/dev/null
cat gpurun_out/r2c16_tables.md | cut -c1-260
