"""Times solver construction phases and host-buffer PCG (development aid)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(90, exit=True)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geometricmultigridpressuresolver_b200 import api, domains as D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cache = f"/tmp/gmg_flip{n}.npz"
ctx = api.Context(0); print("ctx ok", flush=True)
if os.path.exists(cache):
    z = np.load(cache); labels, w, off, levels, dx = z["labels"], [z["w0"], z["w1"], z["w2"]], z["off"], int(z["levels"]), float(z["dx"])
else:
    bl, bw, dx = D.flipsplash_domain(n)
    t = time.perf_counter(); labels, w, off, levels = ctx.buildExpandedDomain(bl, bw); print("buildExpandedDomain ms", (time.perf_counter() - t) * 1e3)
    np.savez(cache, labels=labels, w0=w[0], w1=w[1], w2=w[2], off=off, levels=levels, dx=dx)
hi = [int(off[a]) + n for a in range(3)]; print("inputs ok", flush=True)
b = D.random_rhs(labels, dx, 12345)
x = np.zeros_like(b)
for rep in range(3):
    t0 = time.perf_counter()
    s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=(off, hi), doPrintStats=(rep == 2))
    t1 = time.perf_counter()
    x[...] = 0
    t2 = time.perf_counter()
    xo, it, hist = s.solveGeometricConjugateGradient(x, b, 1e-6, 1000, inplace=True)
    t3 = time.perf_counter()
    s.close()
    t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms (lib {s.setup_ms() if False else 0}) pcg(host) {1e3*(t3-t2):.1f} ms destroy {1e3*(t4-t3):.1f} ms iters {it}")
s = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels)  # no box hint: host scan
print("create without hint: setup_ms", s.setup_ms())
