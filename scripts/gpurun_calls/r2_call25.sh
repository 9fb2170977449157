#!/bin/bash
# round 2, call 25 (= call 20 re-run on the final library) (1 GPU): the whole GPU test-suite, smoke, the default bench line and the ncu launch list on the final library
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c25_pytest.log; tail -5 gpurun_out/r2c25_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c25_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c25_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c25_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "gs", d["gauss_seidel"]["solve_ms"]); print("mixed", d["mixed_precision"]); print("sweep", d["sweep512"]["vcycle_ms"]); print("roofline", d["roofline"]); print("launches", d["gpu_launches"], "clocks", d.get("clocks"))
PY


