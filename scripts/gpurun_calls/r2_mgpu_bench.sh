#!/bin/bash
# round 2, multi-GPU bench line only: bash scripts/gpurun_calls/r2_mgpu_bench.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 297$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b${N}_bench.json 2> gpurun_out/r2b${N}_bench.err; echo "bench rc=$? wall ${SECONDS}s"
tail -3 gpurun_out/r2b${N}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2b${N}_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "parity_vs_n1", d.get("parity_vs_n1"), "failed", d.get("parity_failed"))
s = d.get("sweep512") or {}
print("sweep512", s.get("vcycle_ms"))
b = d.get("solve512") or {}
print("solve512", b.get("solve_ms"), b.get("iterations"), b.get("agreement_1_vs_n"), b.get("halo_exchange_ms_per_solve"))
for k, v in (b.get("fine_level_kernels") or {}).items(): print("   %-16s %8.1f us  frac %.2f share %.2f" % (k, v["us_per_launch"], v["frac_of_hbm_peak"], v["share_of_solve"]))
n = d.get("narrow1024") or {}
print("narrow", n.get("solve_ms"), n.get("agreement_1_vs_n"))
PY
