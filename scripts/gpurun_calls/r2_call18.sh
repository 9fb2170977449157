#!/bin/bash
# round 2, call 18 (1 GPU): full-grid stencil kernels capped at 40 registers (6 resident CTAs per SM instead of 4): A/B per mode
mkdir -p gpurun_out
for v in "GMG_STENCIL_CAP=0" "GMG_STENCIL_CAP=7" "GMG_STENCIL_CAP=1" "GMG_STENCIL_CAP=2" "GMG_STENCIL_CAP=4"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c18_ab_$tag.json 2> gpurun_out/r2c18_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_STENCIL_CAP=0" "GMG_STENCIL_CAP=7"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c18_sweep_$tag.json 2> gpurun_out/r2c18_sweep_$tag.err; echo "sweep $v rc=$?"
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c18_128_$tag.json 2> gpurun_out/r2c18_128_$tag.err; echo "128 $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c18_ab_*.json gpurun_out/r2c18_sweep_*.json gpurun_out/r2c18_128_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|us x"
