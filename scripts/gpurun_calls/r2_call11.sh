#!/bin/bash
# round 2, call 11 (1 GPU): resident band sweep groups (k_band_resident): tests, A/B at 256^3 and 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log; tail -4 gpurun_out/r2c11_pytest.log
for v in "GMG_NONE=1" "GMG_BAND_RESIDENT=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c11_ab_$tag.json 2> gpurun_out/r2c11_ab_$tag.err; echo "$v rc=$?"
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c11_sweep_$tag.json 2> gpurun_out/r2c11_sweep_$tag.err; echo "sweep $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c11_ab_*.json gpurun_out/r2c11_sweep_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|roofline|L[0-9]:|us x"
