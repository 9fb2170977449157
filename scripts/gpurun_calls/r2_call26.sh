#!/bin/bash
# round 2, call 26 (1 GPU): the two new front-end builders (material labels, valid faces) against the restatement; A/B of the one-launch band
# sweep groups restricted to SMALL levels (GMG_BAND_GROUP_MAX = band cells): register-resident groups and grid-barrier groups
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend.py -m gpu -x -q > gpurun_out/r2c26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c26_pytest.log; tail -4 gpurun_out/r2c26_pytest.log
for v in "GMG_BAND_RESIDENT=0" "GMG_BAND_RESIDENT=1 GMG_BAND_GROUP_MAX=40000" "GMG_BAND_RESIDENT=1 GMG_BAND_GROUP_MAX=80000" "GMG_BAND_RESIDENT=1 GMG_BAND_GROUP_MAX=200000" "GMG_BAND_GROUPS=1 GMG_BAND_GROUP_MAX=40000" "GMG_BAND_GROUPS=1 GMG_BAND_GROUP_MAX=200000"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c26_ab_$tag.json 2> gpurun_out/r2c26_ab_$tag.err; echo "$v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c26_ab_*.json 2>/dev/null | grep -E "==|value|vcycle_ms"
