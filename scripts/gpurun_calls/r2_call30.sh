#!/bin/bash
# round 2, call 30 (1 GPU): zero-aware interior sweep rebuilt on a per-cell neighbour-mask byte (one pass, four planes in flight, predicated loads
# near the band) instead of stream-then-patch: bitwise tests, then A/B at 256^3 / 128^3 and on the 512^3 V-cycle
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c30_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c30_pytest.log; tail -4 gpurun_out/r2c30_pytest.log
for v in "GMG_STENCIL_CAP=-1" "GMG_STENCIL_CAP=14" "GMG_ZERO_AWARE=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c30_ab_$tag.json 2> gpurun_out/r2c30_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_STENCIL_CAP=-1" "GMG_STENCIL_CAP=8"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c30_128_$tag.json 2> gpurun_out/r2c30_128_$tag.err; echo "128 $v rc=$?"
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c30_sweep_$tag.json 2> gpurun_out/r2c30_sweep_$tag.err; echo "sweep $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c30_ab_*.json gpurun_out/r2c30_128_*.json gpurun_out/r2c30_sweep_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|L1:|jacobi"
