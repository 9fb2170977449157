"""Communication micro-benchmark of the sharded solver (torchrun, one process per GPU): back-to-back halo exchanges,
gathers and scalar all-reduces, peer-memory kernels vs NCCL (GMG_P2P=0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = api.Context(local)
ctx.shard_with_torch(dist)
bl, bw, dx = D.flipsplash_domain(n)
labels, w, off, levels = ctx.buildExpandedDomain(bl, bw)
s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels)
out = {"world": dist.get_world_size(), "p2p": os.environ.get("GMG_P2P", "1") != "0", "size": n}
for depth in (1, 8):
    for level in (0, 1):
        if s.shard_info(level)[0]:
            out[f"halo_L{level}_d{depth}_us"] = round(s.comm_benchmark(0, level, depth, 200) * 1e3, 2)
out["gather_us"] = round(s.comm_benchmark(1, 0, 0, 200) * 1e3, 2)
out["scalar_us"] = round(s.comm_benchmark(2, 0, 0, 500) * 1e3, 2)
if dist.get_rank() == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
s.close()
ctx.close()
dist.destroy_process_group()
