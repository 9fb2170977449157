"""One V-cycle of the 256^3 flipSplash problem with the tiled Gauss-Seidel smoother between cudaProfilerStart/Stop (for ncu), and
its per-kernel-class event times.  usage: python scripts/profile_gs.py [size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = api.Context(0)
bl, bw, dx = D.flipsplash_domain(n)
labels, w, off, levels = ctx.buildExpandedDomain(bl, bw)
hi = [int(off[a]) + bl.shape[2 - a] for a in range(3)]
s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, useGaussSeidel=True, box=(off, hi))
b = D.random_rhs(labels, dx, 12345)
B, Z = s.grid(0, b), s.grid(0)
for _ in range(2):
    s.applyVCycleDevice(Z, B)
ctx.synchronize()
torch.cuda.profiler.start()
s.applyVCycleDevice(Z, B)
ctx.synchronize()
torch.cuda.profiler.stop()
ts = []
for _ in range(5):
    ctx.timer_begin()
    s.applyVCycleDevice(Z, B)
    ts.append(ctx.timer_end())
print("GS V-cycle ms", np.mean(ts))
ctx.profile_enable(True)
ctx.profile_reset()
s.applyVCycleDevice(Z, B)
for l, row in ctx.profile_by_level(s.getMGLevels()).items():
    print("L%s" % l, {k: (round(v[0], 4), v[1]) for k, v in row.items() if v[1]})
