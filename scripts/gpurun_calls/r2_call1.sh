#!/bin/bash
# round 2, call 1 (1 GPU): GPU test-suite, default bench line, A/B of the round's switches, narrow-band probes, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
for v in "GMG_PDL_PREFETCH=0" "GMG_L2_PERSIST=0" "GMG_DEVICE_LOOP=0" "GMG_PDL_PREFETCH=0 GMG_L2_PERSIST=0 GMG_DEVICE_LOOP=0" "GMG_L2_PERSIST=32" "GMG_PDL=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c1_ab_$tag.json 2> gpurun_out/r2c1_ab_$tag.err; echo "$v rc=$?"
done
timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c1_ab_default.json 2> gpurun_out/r2c1_ab_default.err
timeout 300 python scripts/narrow_probe.py 512 > gpurun_out/r2c1_narrow512.json 2> gpurun_out/r2c1_narrow512.err; echo "narrow512 rc=$?"
timeout 600 python scripts/narrow_probe.py 1024 > gpurun_out/r2c1_narrow1024.json 2> gpurun_out/r2c1_narrow1024.err; echo "narrow1024 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c1_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2c1_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/show_bench.py gpurun_out/r2c1_bench.json 2>/dev/null | head -40
