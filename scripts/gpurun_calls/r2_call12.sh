#!/bin/bash
# round 2, call 12 (1 GPU): mixed precision (fp32 V-cycle inside the fp64 CG), full test-suite, default bench line, ncu traffic for the band kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c12_pytest.log; tail -5 gpurun_out/r2c12_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c12_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c12_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "gs", d["gauss_seidel"]["solve_ms"]); print("mixed", d["mixed_precision"]); print("sweep", d["sweep512"]["vcycle_ms"])
PY
GMG_MIXED=1 timeout 300 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r2c12_sweep_fp64.json 2> /dev/null
timeout 600 ncu --set full --clock-control none -k regex:'k_band|k_stencil' -s 60 -c 14 -o /tmp/pcg256 python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2c12_ncu_pcg.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_table.py /tmp/pcg256.ncu-rep > gpurun_out/r2c12_ncu_pcg256.md 2>&1
GMG_TRAFFIC_JSON=gpurun_out/r2c12_traffic.json python scripts/ncu_traffic.py pcg256=/tmp/pcg256.ncu-rep > /dev/null 2>&1
cat gpurun_out/r2c12_ncu_pcg256.md | cut -c1-220; cat gpurun_out/r2c12_traffic.json
