"""A few V-cycles of the N^3 liquid-box sweep (bench.py --workload vcycle) between cudaProfilerStart/Stop, for ncu.
usage: python scripts/profile_sweep.py [size] [cycles]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = api.Context(0)
bl, bw, dx = D.liquid_box_domain(n)
labels, w, off, levels, box = ctx.buildExpandedDomainLazy(bl, bw)
s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box)
sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
b = np.zeros(labels.shape)
b[sl] = np.random.default_rng(1).random(tuple(x.stop - x.start for x in sl)) * dx * dx * D.active_mask(labels[sl])
B, Z = s.grid(0, b), s.grid(0)
s.applyVCycleDevice(Z, B)
ctx.synchronize()
torch.cuda.profiler.start()
for _ in range(cycles):
    s.applyVCycleDevice(Z, B)
ctx.synchronize()
torch.cuda.profiler.stop()
print("active", s.active_cells(0), "launches", ctx.launch_count())
