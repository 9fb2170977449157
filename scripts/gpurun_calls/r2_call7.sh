#!/bin/bash
# round 2, call 7 (1 GPU): cluster cycle v4 fixed, TMA restriction + prolongation, two-phase zero-aware sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c7_pytest.log; tail -4 gpurun_out/r2c7_pytest.log
for v in "GMG_NONE=1" "GMG_CLUSTER_CYCLE=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c7_ab_$tag.json 2> gpurun_out/r2c7_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_NONE=1" "GMG_TMA=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r2c7_sweep_$tag.json 2> gpurun_out/r2c7_sweep_$tag.err; echo "sweep $v rc=$?"
done
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'k_stencil|k_restrict|k_prolong' -c 8 -o /tmp/sweep512 python scripts/profile_sweep.py 512 1 > gpurun_out/r2c7_ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
python scripts/ncu_table.py /tmp/sweep512.ncu-rep > gpurun_out/r2c7_ncu_sweep512.md 2>&1
cat gpurun_out/r2c7_ncu_sweep512.md | cut -c1-250
python scripts/show_bench.py gpurun_out/r2c7_ab_*.json gpurun_out/r2c7_sweep_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L[0-9]:|us x"
du -sh gpurun_out
