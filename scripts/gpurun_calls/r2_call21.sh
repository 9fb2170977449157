#!/bin/bash
# round 2, call 21 (1 GPU): warp-shuffle taps in the restriction and the full-grid stencil: parity tests, A/B at 256^3 / 128^3 / 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c21_pytest.log; tail -4 gpurun_out/r2c21_pytest.log
for v in "GMG_NONE=1" "GMG_RESTRICT_SHFL=0" "GMG_STENCIL_SHFL=0" "GMG_RESTRICT_SHFL=0 GMG_STENCIL_SHFL=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c21_ab_$tag.json 2> gpurun_out/r2c21_ab_$tag.err; echo "$v rc=$?"
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c21_sweep_$tag.json 2> gpurun_out/r2c21_sweep_$tag.err; echo "sweep $v rc=$?"
done
for v in "GMG_NONE=1" "GMG_RESTRICT_SHFL=0 GMG_STENCIL_SHFL=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c21_128_$tag.json 2> gpurun_out/r2c21_128_$tag.err; echo "128 $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c21_ab_*.json gpurun_out/r2c21_sweep_*.json gpurun_out/r2c21_128_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|L1:|us x"
