#!/bin/bash
# round 2, call 22 (1 GPU): the north_star's 512^3 free-surface solve as a bench block (solve512): stand-alone line, then the default run with it
mkdir -p gpurun_out
echo skip-standalone
SECONDS=0; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c22_bench.json 2> gpurun_out/r2c22_bench.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c22_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "sweep", d["sweep512"]["vcycle_ms"], "solve512", d["solve512"]["solve_ms"], d["solve512"]["iterations"])
PY
