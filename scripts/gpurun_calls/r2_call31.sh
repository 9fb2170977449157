#!/bin/bash
# round 2, call 31 (1 GPU): full-grid stencil kernels with the loads of 2 / 4 planes issued together (k_stencil_b, GMG_STENCIL_BATCH) and the
# branch-free zero-aware sweep: bitwise path test, A/B at 256^3 / 128^3 and on the 512^3 V-cycle (with and without the TMA kernels)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "paths_agree or golden_vcycle or oracle_parity" > gpurun_out/r2c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c31_pytest.log; tail -3 gpurun_out/r2c31_pytest.log
for v in "GMG_STENCIL_BATCH=0" "GMG_STENCIL_BATCH=2" "GMG_STENCIL_BATCH=4"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c31_ab_$tag.json 2> gpurun_out/r2c31_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_STENCIL_BATCH=0" "GMG_STENCIL_BATCH=4"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c31_128_$tag.json 2> gpurun_out/r2c31_128_$tag.err; echo "128 $v rc=$?"
done
for v in "GMG_STENCIL_BATCH=0" "GMG_STENCIL_BATCH=4" "GMG_STENCIL_BATCH=4 GMG_TMA=16" "GMG_STENCIL_BATCH=2 GMG_TMA=16"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 15 --warmup 4 > gpurun_out/r2c31_sweep_$tag.json 2> gpurun_out/r2c31_sweep_$tag.err; echo "sweep $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c31_ab_*.json gpurun_out/r2c31_128_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|L1:"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c31_sweep_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print("==",f, d["value"])
    for k,v in (d.get("fine_level_kernels") or {}).items(): print("     %-16s %8.1f us x%d frac %.3f"%(k,v["us_per_launch"],v["launches_per_vcycle"],v["frac_of_hbm_peak"]))
PY
