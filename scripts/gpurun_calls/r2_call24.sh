#!/bin/bash
# round 2, call 24 (1 GPU): Gauss-Seidel tile kernel with its prologue loads batched: parity tests, per-level event times, GS solve
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c24_pytest.log; tail -4 gpurun_out/r2c24_pytest.log
timeout 300 python scripts/profile_gs.py 256 > gpurun_out/r2c24_gs_events.txt 2>&1; tail -9 gpurun_out/r2c24_gs_events.txt
GMG_GS_V1=1 timeout 300 python scripts/profile_gs.py 256 2>&1 | grep "V-cycle ms"
timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c24_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "gs", d["gauss_seidel"])
PY
