#!/bin/bash
# round 2, call 2 (1 GPU): cluster coarse cycle + zero-aware down-stroke: tests, A/B, default bench line, sanitizer on smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -15 gpurun_out/r2c2_pytest.log
for v in "GMG_NONE=1" "GMG_CLUSTER_CYCLE=0" "GMG_ZERO_AWARE=0" "GMG_CLUSTER_SIZE=8" "GMG_FUSED_FIRST=3"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c2_ab_$tag.json 2> gpurun_out/r2c2_ab_$tag.err; echo "$v rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2c2_memcheck.log gpurun_out/r2c2_racecheck.log
python scripts/show_bench.py gpurun_out/r2c2_ab_GMG_NONE_1.json gpurun_out/r2c2_ab_GMG_CLUSTER_CYCLE_0.json gpurun_out/r2c2_ab_GMG_ZERO_AWARE_0.json 2>/dev/null | grep -E "==|value|vcycle_ms|L[0-9]:"
