"""Ad-hoc GPU-vs-oracle comparison table (development aid; the judged checks are tests/ -m gpu)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bindings import PortLib
from geometricmultigridpressuresolver_b200 import domains as D
from geometricmultigridpressuresolver_b200 import api

def rel(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dom = sys.argv[2] if len(sys.argv) > 2 else "sphere"
    port = PortLib()
    ctx = api.Context(0)
    bl, bw, dx = D.DOMAINS[dom](N)
    L, W, off, lv = port.expand_domain(bl, bw)
    t = time.time(); Lg, Wg, offg, lvg = ctx.buildExpandedDomain(bl, bw); print("gpu expand s", time.time() - t)
    print("labels equal", (L == Lg).all(), "levels", lv, lvg, "off", off, offg, "weights equal", all((W[a] == Wg[a]).all() for a in range(3)))
    cg = ctx.buildCoarseCellLabels(L); cp = port.coarsen_labels(L); print("coarsen equal", (cg == cp).all())
    bc_g = ctx.buildBoundaryCells(L); bc_p = port.boundary_cells(L); print("boundary cells", bc_g.shape, bc_p.shape, bc_g.shape == bc_p.shape and (bc_g == bc_p).all())
    t = time.time(); s = api.GeometricMultigridPoissonSolver(ctx, L, W, lv); print("gpu solver create s", time.time() - t, "setup_ms", s.setup_ms())
    sp = port.solver(L, W, lv, False)
    print("levels", s.getMGLevels(), sp.levels, "coarse n", s.coarse_unknowns(), sp.coarse_unknowns, "active", s.active_cells(0))
    for l in range(s.getMGLevels()):
        lg = s.level_labels(l); lp = sp.level_labels(l)
        bg = s.level_boundary_cells(l); bp = sp.level_boundary_cells(l)
        print(f" level {l}: labels equal {(lg == lp).all()} band {bg.shape[0]} equal {bg.shape == bp.shape and (bg == bp).all()}")
    x = D.random_active(L, 1); b = D.random_active(L, 2)
    X = s.grid(0, x); B = s.grid(0, b); R = s.grid(0)
    print("roundtrip", rel(X.download(), x))
    s.jacobiPoissonSmoother(X, B); print("jacobi", rel(X.download(), port.jacobi(x, b, L, W)))
    X.upload(x); s.boundaryJacobiPoissonSmoother(X, B, 3); print("band x3", rel(X.download(), port.boundary_jacobi(x, b, L, sp.level_boundary_cells(0), 3, W)))
    X.upload(x); s.boundaryJacobiPoissonSmoother(X, B, 1); print("band x1", rel(X.download(), port.boundary_jacobi(x, b, L, sp.level_boundary_cells(0), 1, W)))
    X.upload(x); s.applyPoissonMatrix(R, X); print("apply", rel(R.download(), port.apply(x, L, W)))
    s.computePoissonResidual(R, X, B); print("residual", rel(R.download(), port.residual(x, b, L, W)))
    print("dot", s.dotProduct(X, B), port.dot(x, b, L), "norm2", s.squaredL2Norm(X), port.norm2(x, L), "inf", s.infNorm(X), port.inf_norm(x, L))
    s.addToVector(X, B, 0.3); print("axpy", rel(X.download(), port.axpy(x, b, 0.3, L)))
    X.upload(x); s.addVectors(R, X, B, -0.7); print("add_scaled", rel(R.download(), port.add_scaled(x, b, -0.7, L)))
    X.upload(x); s.scaleVector(X, 1.7); print("scale", rel(X.download(), port.scale(x, 1.7, L)))
    if s.getMGLevels() > 1:
        L1 = sp.level_labels(1)
        X.upload(x); C1 = s.grid(1); s.downsample(C1, X); print("restrict", rel(C1.download(), port.downsample(x, L1, L)))
        xc = D.random_active(L1, 3); C1.upload(xc); X.upload(x); s.upsampleAndAdd(X, C1); print("prolong", rel(X.download(), port.upsample_add(x, xc, L, L1)))
        # level-1 operators (no weights)
        b1 = D.random_active(L1, 4); B1 = s.grid(1, b1); C1.upload(xc)
        s.jacobiPoissonSmoother(C1, B1); print("jacobi L1", rel(C1.download(), port.jacobi(xc, b1, L1)))
        C1.upload(xc); s.boundaryJacobiPoissonSmoother(C1, B1, 3); print("band L1", rel(C1.download(), port.boundary_jacobi(xc, b1, L1, sp.level_boundary_cells(1), 3)))
    v = s.applyVCycle(np.zeros_like(b), b); vp = sp.vcycle(np.zeros_like(b), b); print("vcycle", rel(v, vp))
    v = s.applyVCycle(x, b, True); vp = sp.vcycle(x, b, True); print("vcycle guess", rel(v, vp))
    c = [int(off[0]) + N // 2] * 3
    rhs = D.delta_rhs(L, c, dx)
    if not rhs.any(): rhs = D.random_rhs(L, dx)
    t = time.time(); xg, itg, hg = s.solveGeometricConjugateGradient(np.zeros_like(rhs), rhs, 1e-6, 1000); tg = time.time() - t
    t = time.time(); xp, itp, hp = sp.pcg(np.zeros_like(rhs), rhs, 1e-6, 1000); tp = time.time() - t
    print("pcg iters", itg, itp, "time gpu(e2e) %.4f port %.4f" % (tg, tp))
    n = min(len(hg), len(hp)); print("hist rel", np.abs(hg[:n] - hp[:n]) / hp[:n]); print("x rel", rel(xg, xp))
    print("launches", ctx.launch_count())

main()
