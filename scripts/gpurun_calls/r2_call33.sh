#!/bin/bash
# round 2, call 33 (1 GPU): FINAL library of the round -- the whole GPU test-suite, smoke, the default bench line, the ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c33_pytest.log; tail -4 gpurun_out/r2c33_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c33_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c33_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c33_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c33_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "setup", d["e2e"]["setup_ms"], "gs", d["gauss_seidel"]["solve_ms"], "mixed", d["mixed_precision"]["solve_ms"]); print("sweep", d["sweep512"]["vcycle_ms"], "solve512", d["solve512"].get("solve_ms"), d["solve512"].get("iterations")); print("roofline", d["roofline"]["frac"], d["roofline"]["frac_per_launch_events"], "vcycle", d["vcycle_ms"], "launches", d["gpu_launches"], "clocks", d.get("clocks")); print("parity", d["parity_vs_cpu"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2c33_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2c33_ncu_bench.log 2>&1; echo "ncu rc=$?"
