#!/bin/bash
# round 2, call 8 (1 GPU): where does k_cluster_cycle spend its time (ncu source-level stalls)? longer sweep512 A/B with per-level times
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cluster_cycle' -s 20 -c 1 -o /tmp/cluster python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2c8_ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
python scripts/ncu_table.py /tmp/cluster.ncu-rep > gpurun_out/r2c8_ncu_cluster.md 2>&1
python scripts/ncu_source_top.py /tmp/cluster.ncu-rep 70 > gpurun_out/r2c8_ncu_cluster_source.txt 2>&1
cat gpurun_out/r2c8_ncu_cluster.md | cut -c1-250; head -90 gpurun_out/r2c8_ncu_cluster_source.txt | cut -c1-220
for v in "GMG_NONE=1" "GMG_TMA=0" "GMG_CLUSTER_CYCLE=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 30 --warmup 5 > gpurun_out/r2c8_sweep_$tag.json 2> gpurun_out/r2c8_sweep_$tag.err; echo "sweep $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c8_sweep_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"])
    for l,row in d["kernels_by_level"].items():
        print("   L"+l, {k[:8]:(v["ms_per_vcycle"],v["launches_per_vcycle"]) for k,v in row.items()})
PY
