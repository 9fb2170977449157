"""One MGPCG solve of the bench workload between cudaProfilerStart/Stop, for ncu (--profile-from-start off).
usage: python scripts/profile_step.py [size] [solves]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cache = f"/tmp/gmg_flip{n}.npz"
ctx = api.Context(0)
if os.path.exists(cache):
    z = np.load(cache)
    labels, w, off, levels, dx = z["labels"], [z["w0"], z["w1"], z["w2"]], z["off"], int(z["levels"]), float(z["dx"])
else:
    bl, bw, dx = D.flipsplash_domain(n)
    labels, w, off, levels = ctx.buildExpandedDomain(bl, bw)
    np.savez(cache, labels=labels, w0=w[0], w1=w[1], w2=w[2], off=off, levels=levels, dx=dx)
hi = [int(off[a]) + n for a in range(3)]
s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=(off, hi))
b = D.random_rhs(labels, dx, 12345)
B, X = s.grid(0, b), s.grid(0)
it, hist = s.solveDevice(X, B, 1e-6, 1000)  # warm-up (also instantiates the graphs)
ctx.synchronize()
torch.cuda.profiler.start()
for _ in range(solves):
    X.zero()
    it, hist = s.solveDevice(X, B, 1e-6, 1000)
ctx.synchronize()
torch.cuda.profiler.stop()
print("iterations", it, "final", hist[-1], "active", s.active_cells(0), "launches", ctx.launch_count())
