#!/bin/bash
# round 2, call 9 (1 GPU): which TMA kernels pay in the un-evented 512^3 V-cycle? (GMG_TMA bit mask: 1 Jacobi, 2 residual, 4 apply, 8 restriction, 16 prolongation)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitwise" > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c9_pytest.log; tail -3 gpurun_out/r2c9_pytest.log
for m in 0 31 2 16 18 1 8 19; do
  GMG_TMA=$m timeout 300 python bench.py --workload vcycle --size 512 --steps 30 --warmup 5 > gpurun_out/r2c9_sweep_tma$m.json 2> gpurun_out/r2c9_sweep_tma$m.err; echo "sweep mask $m rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c9_sweep_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],3), {k[:8]:round(v["us_per_launch"]) for k,v in d["fine_level_kernels"].items()})
PY
timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c9_ab_default.json 2> gpurun_out/r2c9_ab_default.err
python scripts/show_bench.py gpurun_out/r2c9_ab_default.json | grep -E "value|vcycle_ms"
