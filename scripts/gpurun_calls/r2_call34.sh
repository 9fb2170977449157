#!/bin/bash
# round 2, call 34 (2 GPUs): the sharded bench line on the FINAL library (after the zero-aware sweep rewrite) without the 512^3 blocks: parity_vs_n1 asserted
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 3 --warmup 3 --no-sweep > gpurun_out/r2c34_bench_n2.json 2> gpurun_out/r2c34_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c34_bench_n2.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "parity_vs_n1", d.get("parity_vs_n1"), "failed", d.get("parity_failed"), "roofline", d["roofline"]["frac"])
PY
