"""Warm-pool constructor timing with page-locked inputs (GMG_PRINT_STATS=1 prints the phases). usage: python scripts/setup_probe2.py [size]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = api.Context(0)
bl, bw, dx = D.flipsplash_domain(n)
labels, w, off, levels = ctx.buildExpandedDomain(bl, bw)
hi = [int(off[a]) + n for a in range(3)]
rt = torch.cuda.cudart()
labels = np.ascontiguousarray(labels, dtype=np.int32)
w = [np.ascontiguousarray(a) for a in w]
for a in [labels] + w:
    assert int(rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0
for k in range(3):
    print(f"--- constructor {k}", flush=True)
    t0 = time.perf_counter()
    s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=(off, hi))
    t1 = time.perf_counter()
    print(f"--- {1e3 * (t1 - t0):.3f} ms", flush=True)
    s.close()
