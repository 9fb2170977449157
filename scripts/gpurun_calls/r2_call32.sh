#!/bin/bash
# round 2, call 32 (1 GPU): register caps of the Jacobi / zero-aware Jacobi kernels (GMG_STENCIL_CAP bits 3, 4, 5: 40 / 48 / 48 registers) at 256^3
mkdir -p gpurun_out
for v in "GMG_STENCIL_CAP=-1" "GMG_STENCIL_CAP=38" "GMG_STENCIL_CAP=22" "GMG_STENCIL_CAP=54" "GMG_STENCIL_CAP=14"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c32_ab_$tag.json 2> gpurun_out/r2c32_ab_$tag.err; echo "$v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c32_ab_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|L1:|L2:"
