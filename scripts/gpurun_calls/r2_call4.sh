#!/bin/bash
# round 2, call 4 (1 GPU): cluster cycle v3 (>= 1024-cell blocks, level 3+), TMA default on big levels; ncu --set full of the 512^3 level-0 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitwise or golden_pcg or golden_vcycle" > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
tail -5 gpurun_out/r2c4_pytest.log
for v in "GMG_NONE=1" "GMG_CLUSTER_CYCLE=0" "GMG_CLUSTER_SIZE=8"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c4_ab_$tag.json 2> gpurun_out/r2c4_ab_$tag.err; echo "$v rc=$?"
done
GMG_PRINT_STATS=1 timeout 120 python -c "
import numpy as np
from geometricmultigridpressuresolver_b200 import api, domains as D
ctx=api.Context(0); bl,bw,dx=D.flipsplash_domain(128); labels,w,off,lv=ctx.buildExpandedDomain(bl,bw)
s=api.GeometricMultigridPoissonSolver(ctx,labels,w,lv,doPrintStats=True); b=D.random_rhs(labels,dx,1); x=s.applyVCycle(np.zeros_like(b),b)
" > gpurun_out/r2c4_print_stats.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_stencil|k_prolong|k_restrict|k_band|k_zero' -c 40 -o gpurun_out/r2c4_sweep512 python scripts/profile_sweep.py 512 1 > gpurun_out/r2c4_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
GMG_TMA=0 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'k_stencil' -c 6 -o gpurun_out/r2c4_sweep512_plain python scripts/profile_sweep.py 512 1 > gpurun_out/r2c4_ncu_sweep_plain.log 2>&1; echo "ncu plain rc=$?"
python scripts/show_bench.py gpurun_out/r2c4_ab_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L[0-9]:|us x"
ls -la gpurun_out/*.ncu-rep
