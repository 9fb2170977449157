#!/bin/bash
# round 2, call 28 (1 GPU): profiling mode 2 (one event pair per band sweep group) -- its test, then the default bench line with the band roofline quoted on it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "profile_group or from_zero" > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c28_pytest.log; grep -E "band_jacobi ms|passed|failed|rc=" gpurun_out/r2c28_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c28_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c28_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "gs", d["gauss_seidel"]["solve_ms"]); print("roofline", d["roofline"])
PY
