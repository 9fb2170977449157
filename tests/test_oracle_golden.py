"""CPU: the plain-C oracle (oracle/gmg_oracle.c) against the fixtures the compiled reference produced."""
import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import domains as D
from tests.common import CASES, FULL_CASES, base_inputs, crop, load_golden, relerr, rhs_for, sha

TOL = 1e-12  # fp64 round-off; the oracle follows the reference's operation order


@pytest.mark.parametrize("name", list(CASES))
def test_inputs_reproduce(name):
    """the committed generators still produce the exact inputs the fixtures were made from"""
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    assert sha(bl.astype(np.int32)) == str(g["base_labels_sha"])
    assert "".join(sha(a) for a in bw) == str(g["base_w_sha"])
    assert dx == float(g["dx"])


@pytest.mark.parametrize("name", list(CASES))
def test_labels_and_boundary_lists_bit_exact(port, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = port.expand_domain(bl, bw)
    assert levels == int(g["mg_levels"]) and (off == g["offset"]).all()
    assert sha(labels.astype(np.int32)) == str(g["labels_sha"])
    assert "".join(sha(a) for a in w) == str(g["weights_sha"])
    assert port.unit_test_boundary_cells(labels, w) and port.unit_test_exterior_cells(labels)
    s = port.solver(labels, w, levels, False)
    assert s.levels == int(g["solver_levels"])
    for l in range(s.levels):
        ll = s.level_labels(l)
        cells = s.level_boundary_cells(l)
        assert sha(ll.astype(np.int32)) == str(g[f"labels_sha_L{l}"])
        assert len(cells) == int(g[f"cells_count_L{l}"])
        assert sha(cells.astype(np.int64)) == str(g[f"cells_sha_L{l}"])
        if l > 0:
            assert port.unit_test_coarsening(ll, s.level_labels(l - 1))
            assert port.unit_test_boundary_cells(ll) and port.unit_test_exterior_cells(ll)
        if f"labels_L{l}" in g:
            assert (ll == g[f"labels_L{l}"]).all()
            assert (cells == g[f"cells_L{l}"]).all()


@pytest.mark.parametrize("name", list(CASES))
def test_pcg_history_and_solution(port, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = port.expand_domain(bl, bw)
    b = rhs_for(labels, off, bl.shape, dx)
    assert sha(b) == str(g["rhs_sha"])
    s = port.solver(labels, w, levels, False)
    x, iters, hist = s.pcg(np.zeros_like(b), b, 1e-6, 1000)
    assert iters == int(g["pcg_iterations"])
    assert len(hist) == len(g["pcg_history"])
    assert np.abs(hist - g["pcg_history"]).max() / g["pcg_history"].max() < 1e-9
    assert (np.abs(hist - g["pcg_history"]) / g["pcg_history"]).max() < 1e-7
    xc = crop(x, off, bl.shape)
    gold = g["pcg_x"]
    if gold.shape != xc.shape:
        xc = xc[::4, ::4, ::4]
    assert relerr(xc, gold) < 1e-9
    assert abs((x * x).sum() - float(g["pcg_x_norm2"])) <= 1e-9 * float(g["pcg_x_norm2"])
    # vector-grid invariant: exactly zero off the active cells
    assert not x[~D.active_mask(labels)].any()


@pytest.mark.parametrize("name", list(CASES))
def test_diagonal_pcg_history_and_solution(port, name):
    """The node's other mode (GFS.cpp:485-618): CG.h's driver with the diagonal preconditioner, capped at 60 iterations."""
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = port.expand_domain(bl, bw)
    b = rhs_for(labels, off, bl.shape, dx)
    s = port.solver(labels, w, levels, False)
    x, iters, hist = s.pcg(np.zeros_like(b), b, 1e-6, 60, diagonal=True)
    assert iters == int(g["dpcg_iterations"]) and len(hist) == len(g["dpcg_history"])
    # relative per iteration; a residual that hit round-off (the delta-rhs fixtures converge to 1e-20) is compared absolutely
    assert (np.abs(hist - g["dpcg_history"]) <= 1e-7 * g["dpcg_history"] + 1e-15).all()
    xc = crop(x, off, bl.shape)
    gold = g["dpcg_x"]
    if gold.shape != xc.shape:
        xc = xc[::4, ::4, ::4]
    assert relerr(xc, gold) < 1e-9


@pytest.mark.parametrize("name", FULL_CASES)
def test_vcycle_and_operators(port, name):
    g = load_golden(name)
    bl, bw, dx = base_inputs(name)
    labels, w, off, levels = port.expand_domain(bl, bw)
    s = port.solver(labels, w, levels, False)
    rb = D.random_rhs(labels, dx, seed=7)
    assert relerr(crop(s.vcycle(np.zeros_like(rb), rb), off, bl.shape), g["vcycle_x"]) < TOL
    x0 = D.random_active(labels, 11, scale=dx * dx)
    assert relerr(crop(s.vcycle(x0, rb, True), off, bl.shape), g["vcycle_guess_x"]) < TOL
    xs, bs = D.random_active(labels, 1), D.random_active(labels, 2)
    cells = s.level_boundary_cells(0)
    c = lambda a: crop(a, off, bl.shape)
    assert relerr(c(port.jacobi(xs, bs, labels, w)), g["op_jacobi"]) < TOL
    assert relerr(c(port.boundary_jacobi(xs, bs, labels, cells, 3, w)), g["op_band3"]) < TOL
    assert relerr(c(port.apply(xs, labels, w)), g["op_apply"]) < TOL
    assert relerr(c(port.residual(xs, bs, labels, w)), g["op_residual"]) < TOL
    assert relerr(c(port.gauss_seidel(xs, bs, labels, 1, 1, w)), g["op_gs_odd_fwd"]) < TOL
    assert relerr(c(port.gauss_seidel(xs, bs, labels, 0, 0, w)), g["op_gs_even_bwd"]) < TOL
    assert abs(port.dot(xs, bs, labels) - float(g["op_dot"])) < 1e-11 * abs(float(g["op_norm2"]))
    assert abs(port.norm2(xs, labels) - float(g["op_norm2"])) < 1e-12 * float(g["op_norm2"])
    assert port.inf_norm(xs, labels) == float(g["op_inf_norm"])
    if s.levels > 1:
        l1 = s.level_labels(1)
        assert relerr(port.downsample(xs, l1, labels), g["op_downsample"]) < TOL
        xc1, b1 = D.random_active(l1, 3), D.random_active(l1, 4)
        assert relerr(c(port.upsample_add(xs, xc1, labels, l1)), g["op_upsample"]) < TOL
        assert relerr(port.jacobi(xc1, b1, l1), g["op_jacobi_L1"]) < TOL
        assert relerr(port.boundary_jacobi(xc1, b1, l1, s.level_boundary_cells(1), 3), g["op_band3_L1"]) < TOL
