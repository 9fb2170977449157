#!/bin/bash
# round 2, multi-GPU call: bash scripts/gpurun_calls/r2_mgpu.sh N   (run under gpurun --gpus N)
# the FULL parity worker (4 domains, S = 1..3, bitwise V-cycle vs the unsharded solver) and the bench line with parity_vs_n1 / sweep512 (/ narrow1024 at N = 8)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2m${N}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N tests/mgpu_worker.py > gpurun_out/r2m${N}_worker.log 2>&1; echo "worker rc=$?" | tee -a gpurun_out/r2m${N}_worker.log
grep -E "SHARD_|FAIL|PCG|levels" gpurun_out/r2m${N}_worker.log | tail -20
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2m${N}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2m${N}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, "parity_vs_n1", d.get("parity_vs_n1"), "failed", d.get("parity_failed"))
    s = d.get("sweep512") or {}
    print("sweep512", s.get("vcycle_ms"), s.get("vcycle_frac_of_hbm_peak"), s.get("halo_exchange_ms_per_vcycle"))
    print("narrow", json.dumps(d.get("narrow1024"))[:900])
    print("halo", d["kernels"].get("halo_exchange"))
except Exception as e:
    print("ERR", e)
PY
