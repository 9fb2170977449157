#!/bin/bash
# round 2, call 3 (1 GPU): cluster cycle v2 (tables in shared memory, local loads as LDS), TMA-staged stencil A/B at 256^3 and 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitwise or golden_pcg or golden_vcycle" > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
tail -15 gpurun_out/r2c3_pytest.log
for v in "GMG_NONE=1" "GMG_CLUSTER_CYCLE=0" "GMG_FUSED_FIRST=3" "GMG_TMA=1" "GMG_CLUSTER_SIZE=8"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c3_ab_$tag.json 2> gpurun_out/r2c3_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_NONE=1" "GMG_TMA=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r2c3_sweep_$tag.json 2> gpurun_out/r2c3_sweep_$tag.err; echo "sweep $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c3_ab_*.json gpurun_out/r2c3_sweep_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L[0-9]:|us x"
