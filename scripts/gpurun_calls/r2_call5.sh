#!/bin/bash
# round 2, call 5 (1 GPU): cluster barrier probe; ncu --set full of the 512^3 level-0 kernels, exported to CSV on the box (the reports themselves are too large to bring back)
mkdir -p gpurun_out
./scripts/cluster_probe > gpurun_out/r2c5_cluster_probe.txt 2>&1; cat gpurun_out/r2c5_cluster_probe.txt
GMG_PRINT_STATS=1 timeout 120 python -c "
import numpy as np
from geometricmultigridpressuresolver_b200 import api, domains as D
ctx=api.Context(0); bl,bw,dx=D.flipsplash_domain(128); labels,w,off,lv=ctx.buildExpandedDomain(bl,bw)
s=api.GeometricMultigridPoissonSolver(ctx,labels,w,lv,doPrintStats=True); b=D.random_rhs(labels,dx,1); x=s.applyVCycle(np.zeros_like(b),b)
" > gpurun_out/r2c5_print_stats.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log; tail -3 gpurun_out/r2c5_pytest.log
timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c5_ab_default.json 2> gpurun_out/r2c5_ab_default.err
GMG_CLUSTER_CYCLE=0 timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c5_ab_nocluster.json 2> gpurun_out/r2c5_ab_nocluster.err
timeout 700 ncu --set full --clock-control none --profile-from-start off -k regex:'k_stencil|k_prolong|k_restrict|k_band' -c 24 -o /tmp/sweep512 python scripts/profile_sweep.py 512 1 > gpurun_out/r2c5_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
python scripts/ncu_table.py /tmp/sweep512.ncu-rep > gpurun_out/r2c5_ncu_sweep512_tma.md 2>&1
GMG_TMA=0 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:'k_stencil' -c 4 -o /tmp/sweep512_plain python scripts/profile_sweep.py 512 1 > gpurun_out/r2c5_ncu_sweep_plain.log 2>&1; echo "ncu plain rc=$?"
python scripts/ncu_table.py /tmp/sweep512_plain.ncu-rep > gpurun_out/r2c5_ncu_sweep512_plain.md 2>&1
cat gpurun_out/r2c5_ncu_sweep512_tma.md gpurun_out/r2c5_ncu_sweep512_plain.md | cut -c1-250
python scripts/show_bench.py gpurun_out/r2c5_ab_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L[0-9]:|us x"
du -sh gpurun_out
