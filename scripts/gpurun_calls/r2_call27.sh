#!/bin/bash
# round 2, call 27 (1 GPU): Gauss-Seidel wavefront trimmed to the fronts that hold an active cell, gmg_pcg_from_zero (e2e without the x0 upload),
# facade zero detection: the whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c27_pytest.log; tail -5 gpurun_out/r2c27_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c27_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c27_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c27_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "gs", d["gauss_seidel"]["solve_ms"]); print("sweep", d["sweep512"]["vcycle_ms"], "solve512", d["solve512"].get("solve_ms")); print("roofline", d["roofline"]["frac"], "launches", d["gpu_launches"], "clocks", d.get("clocks"))
PY
