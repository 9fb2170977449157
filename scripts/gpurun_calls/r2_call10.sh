#!/bin/bash
# round 2, call 10 (1 GPU): full GPU test-suite (Gauss-Seidel v2, front-end builders), default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c10_pytest.log; tail -6 gpurun_out/r2c10_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/r2c10_bench.json | head -30
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c10_bench.json").read().strip().splitlines()[-1])
print("gs", d["gauss_seidel"]); print("e2e", d["e2e"]); print("parity", d["parity_vs_cpu"]); print("sweep", d["sweep512"]["vcycle_ms"], d["sweep512"]["vcycle_frac_of_hbm_peak"])
PY
GMG_GS_V1=1 timeout 300 python - <<'PY' > gpurun_out/r2c10_gs_v1.txt 2>&1
import numpy as np, torch, time
from geometricmultigridpressuresolver_b200 import api, domains as D
ctx=api.Context(0); bl,bw,dx=D.flipsplash_domain(256); labels,w,off,lv=ctx.buildExpandedDomain(bl,bw)
b=D.random_rhs(labels,dx,12345)
s=api.GeometricMultigridPoissonSolver(ctx,labels,w,lv,useGaussSeidel=True); B,X=s.grid(0,b),s.grid(0)
for k in range(4):
    X.zero(); ctx.timer_begin(); it,h=s.solveDevice(X,B,1e-6,1000); ms=ctx.timer_end()
print("GS v1 solve ms", ms, "iterations", it)
PY
cat gpurun_out/r2c10_gs_v1.txt | tail -2
