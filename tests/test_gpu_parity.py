"""GPU (-m gpu): the CUDA path, called through the C ABI (include/gmg_b200.h via the ctypes mirror), against
(a) the fixtures the compiled reference produced and (b) the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): labels and boundary lists bit-exact; per-iteration residual history within
1e-5 relative; iteration count within +-1; final pressure within 1e-5 relative L-inf.  The kernels keep the
reference's operation order, so the asserts below are far tighter than those bars (TOL_*)."""
import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import api
from geometricmultigridpressuresolver_b200 import domains as D
from tests.common import CASES, FULL_CASES, base_inputs, crop, load_golden, relerr, rhs_for, sha

pytestmark = pytest.mark.gpu

TOL_OP = 1e-12       # single operator vs reference, relative L-inf
TOL_HISTORY = 1e-8   # residual history, relative per iteration (bar: 1e-5)
TOL_X = 1e-9         # final pressure, relative L-inf (bar: 1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_golden_labels_and_boundary_lists_bit_exact(gpu_ctx, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    assert levels == int(g["mg_levels"]) and (off == g["offset"]).all()
    assert sha(labels.astype(np.int32)) == str(g["labels_sha"])
    assert "".join(sha(a) for a in w) == str(g["weights_sha"])
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    assert s.getMGLevels() == int(g["solver_levels"])
    for l in range(s.getMGLevels()):
        ll, cells = s.level_labels(l), s.level_boundary_cells(l)
        assert sha(ll.astype(np.int32)) == str(g[f"labels_sha_L{l}"])
        assert len(cells) == int(g[f"cells_count_L{l}"])
        assert sha(cells.astype(np.int64)) == str(g[f"cells_sha_L{l}"])
    # stand-alone builders (buildCoarseCellLabels / buildBoundaryCells on host arrays)
    if "labels_L1" in g:
        assert (gpu_ctx.buildCoarseCellLabels(labels) == g["labels_L1"]).all()
        assert (gpu_ctx.buildBoundaryCells(labels, 3) == g["cells_L0"]).all()
    s.close()


@pytest.mark.parametrize("name", list(CASES))
def test_golden_pcg(gpu_ctx, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = rhs_for(labels, off, bl.shape, dx)
    assert sha(b) == str(g["rhs_sha"])
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    x, iters, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 1000)
    assert abs(iters - int(g["pcg_iterations"])) <= 1 and iters == int(g["pcg_iterations"])
    assert (np.abs(hist - g["pcg_history"]) / g["pcg_history"]).max() < TOL_HISTORY
    xc = crop(x, off, bl.shape)
    gold = g["pcg_x"]
    if gold.shape != xc.shape:
        xc = xc[::4, ::4, ::4]
    assert relerr(xc, gold) < TOL_X
    assert not x[~D.active_mask(labels)].any()
    s.close()


@pytest.mark.parametrize("name", list(CASES))
def test_golden_diagonal_pcg(gpu_ctx, name):
    """preconditioner = 2: the reference node's diagonal-preconditioned mode (GFS.cpp:485-618), 60 iterations of the
    reference's own CG driver on the fixture; also maxIt = 0 (CG.h:100: x untouched, iteration 0)."""
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = rhs_for(labels, off, bl.shape, dx)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    x, iters, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 60, useMGPreconditioner="diagonal")
    assert iters == int(g["dpcg_iterations"]) and len(hist) == len(g["dpcg_history"])
    # relative per iteration; a residual that hit round-off (the delta-rhs fixtures converge to 1e-20) is compared absolutely
    assert (np.abs(hist - g["dpcg_history"]) <= 1e-7 * g["dpcg_history"] + 1e-15).all()
    xc = crop(x, off, bl.shape)
    gold = g["dpcg_x"]
    if gold.shape != xc.shape:
        xc = xc[::4, ::4, ::4]
    assert relerr(xc, gold) < TOL_X
    x0 = D.random_active(labels, 5, scale=dx * dx)
    x1, it1, h1 = s.solveGeometricConjugateGradient(x0, b, 1e-6, 0)
    assert it1 == 0 and len(h1) == 0 and (x1 == x0).all()
    s.close()


def test_byte_labels_and_unit_weights_constructor(gpu_ctx):
    """gmg_solver_create_u8 (one-byte labels) builds the same solver as the int32 form; weights = None is the operators'
    `boundaryWeights == nullptr` form (Ops.h:237-248) = unit weights."""
    bl, bw, dx = base_inputs("narrow_band32")
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = D.random_rhs(labels, dx, seed=3)
    s32 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    s8 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels.astype(np.uint8), w, levels)
    assert s8.getMGLevels() == s32.getMGLevels()
    for l in range(s32.getMGLevels()):
        assert (s8.level_labels(l) == s32.level_labels(l)).all()
        assert (s8.level_boundary_cells(l) == s32.level_boundary_cells(l)).all()
    x32, it32, h32 = s32.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 100)
    x8, it8, h8 = s8.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 100)
    assert it8 == it32 and (h8 == h32).all() and (x8 == x32).all()
    ones = [np.ones_like(a) for a in w]
    s1 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, ones, levels)
    sn = api.GeometricMultigridPoissonSolver(gpu_ctx, labels.astype(np.uint8), None, levels)
    x1, it1, h1 = s1.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 100)
    xn, itn, hn = sn.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 100)
    assert itn == it1 and relerr(hn, h1) < 1e-12 and relerr(xn, x1) < 1e-12
    for s in (s32, s8, s1, sn):
        s.close()


@pytest.mark.parametrize("name", FULL_CASES)
def test_golden_vcycle_and_operators(gpu_ctx, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    c = lambda a: crop(a, off, bl.shape)
    rb = D.random_rhs(labels, dx, seed=7)
    assert relerr(c(s.applyVCycle(np.zeros_like(rb), rb)), g["vcycle_x"]) < TOL_OP
    x0 = D.random_active(labels, 11, scale=dx * dx)
    assert relerr(c(s.applyVCycle(x0, rb, True)), g["vcycle_guess_x"]) < TOL_OP
    xs, bs = D.random_active(labels, 1), D.random_active(labels, 2)
    X, B, R = s.grid(0, xs), s.grid(0, bs), s.grid(0)
    s.jacobiPoissonSmoother(X, B)
    assert relerr(c(X.download()), g["op_jacobi"]) < TOL_OP
    X.upload(xs)
    s.boundaryJacobiPoissonSmoother(X, B, 3)
    assert relerr(c(X.download()), g["op_band3"]) < TOL_OP
    X.upload(xs)
    s.applyPoissonMatrix(R, X)
    assert relerr(c(R.download()), g["op_apply"]) < TOL_OP
    s.computePoissonResidual(R, X, B)
    assert relerr(c(R.download()), g["op_residual"]) < TOL_OP
    assert abs(s.dotProduct(X, B) - float(g["op_dot"])) < 1e-11 * float(g["op_norm2"])
    assert abs(s.squaredL2Norm(X) - float(g["op_norm2"])) < 1e-12 * float(g["op_norm2"])
    assert s.infNorm(X) == float(g["op_inf_norm"])
    if s.getMGLevels() > 1:
        l1 = s.level_labels(1)
        C1 = s.grid(1)
        s.downsample(C1, X)
        assert relerr(C1.download(), g["op_downsample"]) < TOL_OP
        xc1, b1 = D.random_active(l1, 3), D.random_active(l1, 4)
        C1.upload(xc1)
        s.upsampleAndAdd(X, C1)
        assert relerr(c(X.download()), g["op_upsample"]) < TOL_OP
        B1 = s.grid(1, b1)
        s.jacobiPoissonSmoother(C1, B1)
        assert relerr(C1.download(), g["op_jacobi_L1"]) < TOL_OP
        C1.upload(xc1)
        s.boundaryJacobiPoissonSmoother(C1, B1, 3)
        assert relerr(C1.download(), g["op_band3_L1"]) < TOL_OP
    s.close()


# ---- against the CPU oracle on the BASELINE configs that fit a test (SURVEY.md 8d configs 1, 2 and small 3/4/5) ----
ORACLE_CASES = [("sphere", 64, {}), ("sphere", 128, {}), ("flipsplash", 96, {"shape": (96, 64, 96)}), ("liquid_box", 64, {}),
                ("narrow_band", 96, {}), ("complex", 64, {}), ("simple", 48, {})]


@pytest.mark.parametrize("dom,n,kw", ORACLE_CASES)
def test_oracle_parity_full_solve(gpu_ctx, port, dom, n, kw):
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    pl, pw, poff, plv = port.expand_domain(bl, bw)
    assert levels == plv and (labels == pl).all() and all((w[a] == pw[a]).all() for a in range(3))
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    ps = port.solver(pl, pw, plv, False)
    assert s.getMGLevels() == ps.levels and s.coarse_unknowns() == ps.coarse_unknowns
    for l in range(ps.levels):
        assert (s.level_labels(l) == ps.level_labels(l)).all()
        gc, pc = s.level_boundary_cells(l), ps.level_boundary_cells(l)
        assert gc.shape == pc.shape and (gc == pc).all()
    b = rhs_for(labels, off, bl.shape, dx)
    x, it, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 1000)
    xo, ito, histo = ps.pcg(np.zeros_like(b), b, 1e-6, 1000)
    assert abs(it - ito) <= 1 and it == ito
    assert (np.abs(hist - histo) / histo).max() < TOL_HISTORY
    assert relerr(x, xo) < TOL_X
    # warm start (production passes the old pressure, GFS.cpp:408-418)
    x2, it2, hist2 = s.solveGeometricConjugateGradient(0.5 * xo, b, 1e-6, 1000)
    xo2, ito2, histo2 = ps.pcg(0.5 * xo, b, 1e-6, 1000)
    assert it2 == ito2 and (np.abs(hist2 - histo2) / histo2).max() < 1e-6 and relerr(x2, xo2) < TOL_X
    s.close()


def test_early_outs_and_errors(gpu_ctx):
    bl, bw, dx = D.sphere_domain(24)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    zero = np.zeros(labels.shape)
    x, it, hist = s.solveGeometricConjugateGradient(zero, zero, 1e-6, 10)  # "RHS is zero" (CG.h:35-40)
    assert it == -1 and len(hist) == 0 and not x.any()
    b = rhs_for(labels, off, bl.shape, dx)
    xs, it, hist = s.solveGeometricConjugateGradient(zero, b, 1e-10, 1000)
    x2, it2, hist2 = s.solveGeometricConjugateGradient(xs, b, 1e-6, 1000)  # "Residual already below error" (CG.h:60-64)
    assert it2 == -1 and len(hist2) == 0 and (x2 == xs).all()
    x3, it3, hist3 = s.solveGeometricConjugateGradient(zero, b, 1e-12, 3)  # iteration cap: CG.h:198 prints maxIterations
    assert it3 == 3 and len(hist3) == 3
    with pytest.raises(api.GmgError):  # odd resolution (MG.cpp:155-157)
        api.GeometricMultigridPoissonSolver(gpu_ctx, labels[:-1], [w[0][:-1], w[1][:-1], w[2][:-1]], levels)
    # pure-Neumann box: singular coarse matrix is reported, not silently "solved"
    box = np.full((16, 16, 16), D.EXTERIOR, dtype=np.int32)
    box[1:-1, 1:-1, 1:-1] = D.INTERIOR
    wbox = []
    for axis in range(3):
        ww = np.zeros(D.face_shape(box.shape, axis))
        sl = [slice(1, -1)] * 3
        sl[2 - axis] = slice(2, -2)
        ww[tuple(sl)] = 1.0
        wbox.append(ww)
    el, ew, eo, elv = gpu_ctx.buildExpandedDomain(box, wbox)
    with pytest.raises(api.GmgError) as e:
        api.GeometricMultigridPoissonSolver(gpu_ctx, el, ew, elv)
    assert e.value.status == 5
    s.close()


def test_pcg_from_zero_equals_pcg_on_a_zero_grid(gpu_ctx):
    """gmg_pcg_from_zero (the caller declares the solution grid constant zero: nothing of it is uploaded) against gmg_pcg on an explicit
    zero array: bitwise the same pressure, history and iteration count; the output array's content on entry does not matter."""
    bl, bw, dx = D.flipsplash_domain(32)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    b = D.random_rhs(labels, dx)
    x1, it1, h1 = s.solveGeometricConjugateGradient(np.zeros(labels.shape), b, 1e-8, 200)
    x2, it2, h2 = s.solveGeometricConjugateGradient(None, b, 1e-8, 200, solutionIsZero=True)
    assert it1 == it2 and (h1 == h2).all() and (x1 == x2).all()
    for pre in ("diagonal", False):
        y1, j1, g1 = s.solveGeometricConjugateGradient(np.zeros(labels.shape), b, 1e-4, 50, useMGPreconditioner=pre)
        y2, j2, g2 = s.solveGeometricConjugateGradient(None, b, 1e-4, 50, useMGPreconditioner=pre, solutionIsZero=True)
        assert j1 == j2 and (g1 == g2).all() and (y1 == y2).all()
    x3, it3, h3 = s.solveGeometricConjugateGradient(np.zeros(labels.shape), np.zeros(labels.shape), 1e-6, 10, inplace=True, solutionIsZero=True)
    assert it3 == -1 and not x3.any()  # "RHS is zero" (CG.h:35-40)
    with pytest.raises(ValueError):
        s.solveGeometricConjugateGradient(None, b, 1e-8, 200)
    s.close()


def test_profile_group_brackets_count_the_same_launches(gpu_ctx):
    """gmg_profile_enable(ctx, 2) puts one event pair around a band sweep group instead of one per sweep: same launches, same
    algorithmic bytes, same result as mode 1; the brackets are fewer, so the class cannot take longer than with a pair per launch."""
    bl, bw, dx = D.flipsplash_domain(48)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    b = D.random_rhs(labels, dx)
    B = s.grid(0, b)
    prof, sol = {}, {}
    for mode in (1, 2):
        X = s.grid(0)
        s.solveDevice(X, B, 1e-6, 100)  # captures the profiled graph of this mode
        gpu_ctx.profile_enable(mode)
        gpu_ctx.profile_reset()
        X.zero()
        it, hist = s.solveDevice(X, B, 1e-6, 100)[:2]
        prof[mode] = gpu_ctx.profile(False)["band_jacobi"]
        sol[mode] = (it, X.download())
        gpu_ctx.profile_enable(False)
    assert prof[1][1] == prof[2][1] > 0 and prof[1][2] == prof[2][2]
    print("band_jacobi ms: a pair per launch", prof[1][0], "a pair per group", prof[2][0])
    assert prof[2][0] < 1.05 * prof[1][0]
    assert sol[1][0] == sol[2][0] and (sol[1][1] == sol[2][1]).all()
    s.close()


def test_diagonal_free_plain_cg_matches_oracle_operators(gpu_ctx, port):
    """preconditioner = 0 runs plain CG on the same kernels; cross-check with numpy CG over oracle operators."""
    bl, bw, dx = D.complex_domain(24)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    b = D.random_rhs(labels, dx, 3)
    # unpreconditioned CG on this system is chaotic: two summation orders drift apart by ~2.4x per iteration (1e-15 at
    # iteration 3, O(1) by iteration 40 -- measured), so only a short prefix can be compared
    cap = 12
    x, it, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, cap, useMGPreconditioner=False)
    xr = np.zeros_like(b)
    r = b.copy()
    p = r.copy()
    rho = port.dot(p, r, labels)
    bb = port.norm2(b, labels)
    hist_o = []
    for k in range(cap):
        t = port.apply(p, labels, w)
        alpha = rho / port.dot(p, t, labels)
        xr = port.axpy(xr, p, alpha, labels)
        r = port.axpy(r, t, -alpha, labels)
        rr = port.norm2(r, labels)
        hist_o.append(np.sqrt(rr / bb))
        if rr < 1e-12 * bb:
            break
        rho_new = port.dot(r, r, labels)
        p = port.add_scaled(r, p, rho_new / rho, labels)
        rho = rho_new
    assert it == cap and len(hist) == len(hist_o) == cap
    assert (np.abs(hist - np.array(hist_o)) / np.array(hist_o)).max() < 1e-8
    assert relerr(x, xr) < 1e-8
    s.close()


def test_coarse_matrix_scale_quirk(gpu_ctx, port):
    bl, bw, dx = D.sphere_domain(20)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = D.random_rhs(labels, dx, 5)
    s3 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels, coarse_matrix_scale=3.0)
    v3 = s3.applyVCycle(np.zeros_like(b), b)
    vo = port.solver(labels, w, levels, False, coarse_scale=3.0).vcycle(np.zeros_like(b), b)
    assert relerr(v3, vo) < 1e-11
    s3.close()


# ---- tiled Gauss-Seidel, the reference's production smoother ----------------------------------------
@pytest.mark.parametrize("dom,n", [("complex", 24), ("flipsplash", 32), ("sphere", 48), ("narrow_band", 32)])
def test_tiled_gauss_seidel_half_passes_match_oracle(gpu_ctx, port, dom, n):
    """Ops.h:369-520: every (odd/even tiles, forwards/backwards) half-pass, at level 0 (face weights) and level 1."""
    bl, bw, dx = D.DOMAINS[dom](n)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels, useGaussSeidel=True)
    for level in range(min(2, s.getMGLevels())):
        ll = s.level_labels(level)
        x0, b0 = D.random_active(ll, 3 + level), D.random_active(ll, 4 + level)
        for odd in (True, False):
            for fwd in (True, False):
                X, B = s.grid(level, x0), s.grid(level, b0)
                s.tiledGaussSeidelPoissonSmoother(X, B, odd, fwd)
                ref = port.gauss_seidel(x0, b0, ll, odd, fwd, w if level == 0 else None)
                assert relerr(X.download(), ref) < TOL_OP, (level, odd, fwd)
    s.close()


@pytest.mark.parametrize("dom,n", [("complex", 32), ("flipsplash", 32), ("liquid_box", 32)])
def test_gauss_seidel_vcycle_and_pcg_match_oracle(gpu_ctx, port, dom, n):
    """useGaussSeidel = true: MG.cpp:466-479 (odd then even tiles, forwards) and :740-751 (even then odd, backwards)."""
    bl, bw, dx = D.DOMAINS[dom](n)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = D.random_rhs(labels, dx, 9)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels, useGaussSeidel=True)
    ps = port.solver(labels, w, levels, True)
    assert s.getMGLevels() == ps.levels
    assert relerr(s.applyVCycle(np.zeros_like(b), b), ps.vcycle(np.zeros_like(b), b)) < 1e-11
    x0 = D.random_active(labels, 2, 1e-3)
    assert relerr(s.applyVCycle(x0, b, useInitialGuess=True), ps.vcycle(x0, b, True)) < 1e-11
    x, it, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 500)
    xo, ito, histo = ps.pcg(np.zeros_like(b), b, 1e-6, 500)
    assert it == ito
    assert (np.abs(hist - histo) / histo).max() < TOL_HISTORY
    assert relerr(x, xo) < TOL_X
    # Gauss-Seidel halves the iteration count of the Jacobi smoother or better (why production uses it)
    sj = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    _, itj, _ = sj.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 500)
    assert it <= itj
    s.close()
    sj.close()


def test_lazy_expanded_domain_matches_the_full_one(gpu_ctx):
    """buildExpandedDomainLazy only writes the base box of the (virtual) expanded grids; with the box hint the library never
    reads outside it, so the solver it produces is the one the fully materialised arrays give."""
    bl, bw, dx = D.flipsplash_domain(32)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    ll, lw, loff, llevels, box = gpu_ctx.buildExpandedDomainLazy(bl, bw)
    assert llevels == levels and (loff == off).all()
    sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
    assert (ll[sl] == labels[sl]).all()
    outside = np.ones(labels.shape, dtype=bool)
    outside[sl] = False
    assert not ll[outside].any()  # zeros, i.e. NOT the EXTERIOR label: only usable together with the box hint
    s1 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    s2 = api.GeometricMultigridPoissonSolver(gpu_ctx, ll, lw, llevels, box=box)
    assert s1.getMGLevels() == s2.getMGLevels()
    for l in range(s1.getMGLevels()):
        assert (s1.level_labels(l) == s2.level_labels(l)).all()
        assert (s1.level_boundary_cells(l) == s2.level_boundary_cells(l)).all()
    b = D.random_rhs(labels, dx, 21)
    x1, it1, h1 = s1.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 200)
    x2, it2, h2 = s2.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 200)
    assert it1 == it2 and np.array_equal(h1, h2) and np.array_equal(x1, x2)
    s1.close()
    s2.close()


@pytest.mark.parametrize("dom,n,kw", [("complex", 32, {}), ("flipsplash", 48, {"shape": (48, 32, 48)}), ("narrow_band", 48, {})])
def test_band_smoother_any_sweep_count_matches_oracle(gpu_ctx, port, dom, n, kw):
    """The temporally blocked band kernel runs groups of up to three sweeps per launch on shrinking shared-memory boxes;
    every sweep count must give what that many separate boundaryJacobiPoissonSmoother calls give (Ops.h:524-619)."""
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    for level in range(min(2, s.getMGLevels())):
        ll = s.level_labels(level)
        cells = s.level_boundary_cells(level)
        lw = w if level == 0 else None
        xs, bs = D.random_active(ll, 31 + level), D.random_active(ll, 41 + level)
        X, B = s.grid(level, xs), s.grid(level, bs)
        for sweeps in (1, 2, 3, 4, 5, 7):
            X.upload(xs)
            s.boundaryJacobiPoissonSmoother(X, B, sweeps)
            want = port.boundary_jacobi(xs.copy(), bs, ll, cells, sweeps, lw)
            assert relerr(X.download(), want) < TOL_OP, (level, sweeps)
    s.close()


def test_pinned_host_buffers_move_only_the_active_rectangles(gpu_ctx):
    """With page-locked caller buffers rhs / x0 / pressure travel as per-z-plane bounding rectangles of the active cells
    (gmg_solver_transfer_cells); the result must be the one the pageable path (whole box through staging buffers) gives,
    and cells outside the rectangles must stay untouched on the host."""
    import torch

    bl, bw, dx = D.flipsplash_domain(48)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    cells, copies = s.transfer_cells()
    box = np.argwhere(labels != 1)
    box_cells = int(np.prod(box.max(0) - box.min(0) + 1))
    assert D.active_mask(labels).sum() <= cells < box_cells and copies >= 1
    b = D.random_rhs(labels, dx, 5)
    x0 = 0.25 * D.random_active(labels, 6, scale=dx * dx)
    x_ref, it_ref, h_ref = s.solveGeometricConjugateGradient(x0, b, 1e-6, 200)
    rt = torch.cuda.cudart()
    xp, bp = np.ascontiguousarray(x0.copy()), np.ascontiguousarray(b.copy())
    for a in (xp, bp):
        assert int(rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0
    try:
        x_pin, it_pin, h_pin = s.solveGeometricConjugateGradient(xp, bp, 1e-6, 200, inplace=True)
        assert it_pin == it_ref and np.array_equal(h_pin, h_ref) and np.array_equal(x_pin, x_ref)
        v_ref = s.applyVCycle(np.zeros_like(b), b)
        xp[...] = 0.0
        v_pin = s.applyVCycle(xp, bp, inplace=True)
        assert np.array_equal(v_pin, v_ref)
    finally:
        for a in (xp, bp):
            rt.cudaHostUnregister(a.ctypes.data)
    s.close()


def test_vcycle_paths_agree_bitwise(gpu_ctx, monkeypatch):
    """The coarse levels have three implementations -- one kernel per operator, the one-CTA shared-memory cycle, the
    thread-block-cluster cycle with distributed shared memory --, the down-stroke two (zero fill + plain Jacobi, or the
    zero-aware sweep) and the full-grid stencil two (plain loads, or TMA-staged bricks).  Per cell they run the same
    arithmetic in the same order: a V-cycle must agree BITWISE between them."""
    bl, bw, dx = D.flipsplash_domain(64)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    rb = D.random_rhs(labels, dx, seed=21)
    x0 = D.random_active(labels, 22, scale=dx * dx)
    results = {}
    for name, env in [("cluster", {"GMG_CLUSTER_CYCLE": "1"}), ("compact", {}), ("kernels", {"GMG_COARSE_FUSED": "0", "GMG_CLUSTER_SMOOTH": "0", "GMG_BAND_TILES": "0", "GMG_BAND_PER_THREAD": "2"}), ("level_kernels", {"GMG_CLUSTER_SMOOTH": "0"}),
                      ("zero_fill", {"GMG_ZERO_AWARE": "0"}), ("cluster8", {"GMG_CLUSTER_CYCLE": "1", "GMG_CLUSTER_SIZE": "8"}), ("first2", {"GMG_CLUSTER_CYCLE": "1", "GMG_FUSED_FIRST": "2"}),
                      ("tma", {"GMG_TMA": "31", "GMG_TMA_MIN_CELLS": "100"}), ("sweep_groups", {"GMG_BAND_GROUPS": "1", "GMG_BAND_TILES": "0", "GMG_BAND_PER_THREAD": "2"}), ("sweep_resident", {"GMG_BAND_RESIDENT": "1", "GMG_BAND_TILES": "0", "GMG_BAND_PER_THREAD": "2"}), ("sweep_kernels", {"GMG_BAND_TILES": "0", "GMG_BAND_PER_THREAD": "2"}),
                      ("stencil_loop", {"GMG_STENCIL_LOOP": "15", "GMG_STENCIL_LOOP_SLOTS": "24"}), ("stencil_cap", {"GMG_STENCIL_CAP": "15"}), ("stencil_cap48", {"GMG_STENCIL_CAP": "54"}),
                      ("stencil_batch2", {"GMG_STENCIL_BATCH": "2"}), ("stencil_batch4", {"GMG_STENCIL_BATCH": "4"}), ("stencil_one_plane", {"GMG_STENCIL_BATCH": "0"})]:
        for k in ("GMG_CLUSTER_CYCLE", "GMG_COARSE_FUSED", "GMG_ZERO_AWARE", "GMG_CLUSTER_SIZE", "GMG_FUSED_FIRST", "GMG_TMA", "GMG_TMA_MIN_CELLS", "GMG_BAND_GROUPS", "GMG_BAND_RESIDENT", "GMG_CLUSTER_SMOOTH", "GMG_BAND_TILES", "GMG_BAND_PER_THREAD", "GMG_STENCIL_LOOP", "GMG_STENCIL_LOOP_SLOTS", "GMG_STENCIL_CAP", "GMG_STENCIL_BATCH"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
        results[name] = (s.applyVCycle(np.zeros_like(rb), rb), s.applyVCycle(x0, rb, True), s.solveGeometricConjugateGradient(np.zeros_like(rb), rb, 1e-8, 50))
        s.close()
    ref = results["kernels"]
    for name, (v, vg, (x, it, hist)) in results.items():
        assert (v == ref[0]).all(), name
        assert (vg == ref[1]).all(), name
        if name in ("tma", "stencil_loop"):
            # the brick kernel sums p.Ap over a different CTA decomposition: same cells, another order of the partial sums
            assert it == ref[2][1] and relerr(hist, ref[2][2]) < 1e-11 and relerr(x, ref[2][0]) < 1e-11, name
        else:
            assert it == ref[2][1] and (hist == ref[2][2]).all() and (x == ref[2][0]).all(), name


@pytest.mark.parametrize("dom,n,kw", [("sphere", 64, {}), ("complex", 48, {}), ("narrow_band", 64, {"thickness": 6}), ("flipsplash", 48, {"shape": (48, 32, 64)}),
                                      ("sphere", 128, {})])
def test_band_sweep_tiles_agree_bitwise(gpu_ctx, monkeypatch, dom, n, kw):
    """A group of three band sweeps runs as three launches over the whole band list (two or three cells per thread) or -- opt-in,
    measured slower -- as ONE launch of ring-halo tiles (gmg_band_tiles.cuh: sweep 1 on a tile + two rings, sweep 2 on the tile + one
    ring, sweep 3 on the tile).  Same per-cell
    arithmetic: the band smoother alone, a V-cycle (zero guess and initial guess) and a PCG solve must agree BITWISE, on weighted
    (ghost-fluid) and unweighted levels, with the tiles used for every group or only for the groups that start from a zero grid."""
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    rb = D.random_rhs(labels, dx, seed=31)
    x0 = D.random_active(labels, 32, scale=dx * dx)
    out = {}
    for name, env in [("sweeps", {"GMG_BAND_PER_THREAD": "2"}), ("tiles", {"GMG_BAND_TILES": "1"}), ("zero_only", {"GMG_BAND_TILES": "1", "GMG_BAND_TILES_ZERO_ONLY": "1"}),
                      ("three_per_thread", {"GMG_BAND_PER_THREAD": "3"}), ("default", {})]:
        for k in ("GMG_BAND_TILES", "GMG_BAND_TILES_ZERO_ONLY", "GMG_BAND_PER_THREAD"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
        X, B = s.grid(0, x0), s.grid(0, rb)
        s.boundaryJacobiPoissonSmoother(X, B, 3)
        out[name] = (X.download(), s.applyVCycle(np.zeros_like(rb), rb), s.applyVCycle(x0, rb, True),
                     s.solveGeometricConjugateGradient(np.zeros_like(rb), rb, 1e-8, 40))
        s.close()
    ref = out["sweeps"]
    for name, (bj, v, vg, (x, it, hist)) in out.items():
        assert (bj == ref[0]).all() and (v == ref[1]).all() and (vg == ref[2]).all(), name
        assert it == ref[3][1] and (hist == ref[3][2]).all() and (x == ref[3][0]).all(), name


@pytest.mark.parametrize("dom,n,kw", [("sphere", 64, {}), ("flipsplash", 64, {}), ("complex", 48, {})])
def test_mixed_precision_preconditioner(gpu_ctx, dom, n, kw):
    """mixed_precision = 1: the V-cycle in fp32 inside the fp64 CG.  Not the reference arithmetic -- what is asserted is
    that the fp64 CG still converges to the same pressure: iteration count within +-2 of the fp64 preconditioner, residual history
    within 5 % per iteration, final pressure within 1e-5 relative of the fp64 run (the solve tolerance is 1e-6), and that the fp64
    path of the same library is untouched (bitwise equal to a solver created without the option)."""
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, levels = gpu_ctx.buildExpandedDomain(bl, bw)
    b = rhs_for(labels, off, bl.shape, dx)
    s64 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels)
    s32 = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels, mixed_precision=True)
    x64, it64, h64 = s64.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 1000)
    x32, it32, h32 = s32.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 1000)
    assert abs(it32 - it64) <= 2 and h32[-1] < 1e-6
    m = min(len(h32), len(h64))
    assert (np.abs(h32[:m] - h64[:m]) / h64[:m]).max() < 0.05
    assert relerr(x32, x64) < 1e-5
    assert not x32[~D.active_mask(labels)].any()
    # the V-cycle entry point of the mixed solver is still the fp64 one
    rb = D.random_rhs(labels, dx, seed=7)
    assert (s32.applyVCycle(np.zeros_like(rb), rb) == s64.applyVCycle(np.zeros_like(rb), rb)).all()
    s64.close()
    s32.close()
