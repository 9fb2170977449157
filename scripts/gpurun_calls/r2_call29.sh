#!/bin/bash
# round 2, call 29 (2 GPUs): the sharded bench line on the final library without the 512^3 blocks -- value, both profiling passes, e2e through
# gmg_pcg_from_zero on two ranks, parity_vs_n1 asserted
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 5 --warmup 3 --no-sweep > gpurun_out/r2c29_bench_n2.json 2> gpurun_out/r2c29_bench_n2.err; echo "bench rc=$? wall ${SECONDS}s"
tail -3 gpurun_out/r2c29_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c29_bench_n2.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"], "parity_vs_n1", d.get("parity_vs_n1"), "failed", d.get("parity_failed"))
print("roofline", d["roofline"]["frac"], d["roofline"].get("frac_per_launch_events"), "launches", d["gpu_launches"], "nccl_ops", d["nccl_ops"])
PY
