"""The reference's own DIAGNOSTIC node -- HDK_TestGeometricMultigrid.cpp, compiled unmodified over oracle/shim into
oracle/_ref/libgmg_ref_testnode.so -- run as its author runs it inside Houdini (the reference has no other tests).

  * its symmetry battery (Test.cpp:1165-1875: Jacobi / Gauss-Seidel smoothing, direct solve, restriction-prolongation, one-level and full
    V-cycle with both smoothers) on both of its domains: the oracle build -- the reference's operator sources over the HDK / Eigen stand-ins --
    passes the reference's own tests;
  * its MGPCG test (Test.cpp:675-1010: its own buildSimpleDomain / buildComplexDomain, its own delta right-hand side, tiled Gauss-Seidel
    V-cycle as the preconditioner) against the C restatement on domains.py's re-creation of those domains: same iteration count, the printed
    residual history to its ten digits.  This pins the synthetic input generators every other test builds on.
  * its smoother test (Test.cpp:1962-2105: band sweeps around a damped-Jacobi or a four-half-pass tiled Gauss-Seidel interior sweep) against the
    same sequence of restated operators: the printed residual norms, round by round.
  * its V-cycle convergence test (Test.cpp:1877-1960: fifty damped-Jacobi V-cycles with useInitialGuess from a sinusoid, zero right-hand side)
    and its diagonally preconditioned CG branch, likewise.
  * tests/golden/reference_testnode.json keeps what the node printed, for the boxes without /root/reference (CPU: the restatement; GPU: the
    CUDA path in Gauss-Seidel mode)."""
import json
import os
import re

import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import domains as D
from tests.common import GOLDEN_DIR

CG_CASES = [("simple", 16), ("simple", 32), ("complex", 16), ("complex", 32), ("complex", 48)]
TOL, MAX_IT, AMPLITUDE = 1e-6, 1000, 1000.0
FIXTURE = os.path.join(GOLDEN_DIR, "reference_testnode.json")


def node_cg(testnode, dom, n):
    ok, log = testnode.run(gridSize=n, useComplexDomain=int(dom == "complex"), testConjugateGradient=1, useMultigridPreconditioner=1, solveCGGeometrically=1,
                           solverTolerance=TOL, maxSolverIterations=MAX_IT, deltaFunctionAmplitude=AMPLITUDE)
    assert ok, log[-2000:]
    hist = [float(x) for x in re.findall(r"Relative error: ([-+.\deE]+)", log)]
    return int(re.findall(r"Iterations: (\d+)", log)[-1]), hist


def delta_problem(expand_domain, dom, n):
    """domains.py's re-creation of the node's domain and of its delta right-hand side (Test.cpp:727-742: the 3^3 block around 10 % of the grid)."""
    bl, bw, dx = D.DOMAINS[dom](n)
    labels, w, off, levels = expand_domain(bl, bw)
    centre = [int(np.float32(0.1) * np.float32(n)) + int(off[a]) for a in range(3)]
    return labels, w, levels, D.delta_rhs(labels, centre, dx, AMPLITUDE)


def parse_symmetry(log):
    rows = re.findall(r"^\s*(.*?)\. BMA: ([-+.\deE]+)\. AMB: ([-+.\deE]+)", log, flags=re.M)
    return [(name.strip(), float(a), float(b)) for name, a, b in rows]


@pytest.mark.parametrize("complex_domain", [0, 1])
def test_the_reference_symmetry_battery_passes_on_the_oracle_build(testnode, complex_domain):
    ok, log = testnode.run(gridSize=16, useComplexDomain=complex_domain, testSymmetry=1)
    assert ok, log[-2000:]
    rows = parse_symmetry(log)
    assert len(rows) == 7, log[-3000:]
    for name, bma, amb in rows:
        assert bma != 0 and abs(bma - amb) <= 1e-9 * abs(bma), (name, bma, amb)


@pytest.mark.parametrize("dom,n", CG_CASES)
def test_the_reference_cg_test_against_the_restatement(testnode, port, dom, n):
    it_node, hist_node = node_cg(testnode, dom, n)
    labels, w, levels, b = delta_problem(port.expand_domain, dom, n)
    x, it, hist = port.solver(labels, w, levels, True).pcg(np.zeros_like(b), b, TOL, MAX_IT)
    assert it == it_node and len(hist) == len(hist_node)
    assert max(abs(a - c) / c for a, c in zip(hist_node, hist)) < 2e-9  # the node prints ten significant digits
    stored = json.load(open(FIXTURE))[f"{dom}{n}"]
    assert stored["iterations"] == it_node and stored["history"] == hist_node, "tests/golden/reference_testnode.json is stale: python tests/golden/make_golden_frontend.py"


@pytest.mark.parametrize("dom,n", CG_CASES)
def test_restatement_against_the_stored_log_of_the_reference_cg_test(port, dom, n):
    stored = json.load(open(FIXTURE))[f"{dom}{n}"]
    labels, w, levels, b = delta_problem(port.expand_domain, dom, n)
    x, it, hist = port.solver(labels, w, levels, True).pcg(np.zeros_like(b), b, TOL, MAX_IT)
    assert it == stored["iterations"] and len(hist) == len(stored["history"])
    assert max(abs(a - c) / c for a, c in zip(stored["history"], hist)) < 2e-9


@pytest.mark.gpu
@pytest.mark.parametrize("dom,n", [("simple", 32), ("complex", 32), ("complex", 48)])
def test_gpu_against_the_stored_log_of_the_reference_cg_test(gpu_ctx, dom, n):
    """The CUDA path in the node's configuration (useGaussSeidel = true, Test.cpp:805-808) on the node's own problem."""
    from geometricmultigridpressuresolver_b200 import api

    stored = json.load(open(FIXTURE))[f"{dom}{n}"]
    labels, w, levels, b = delta_problem(gpu_ctx.buildExpandedDomain, dom, n)
    s = api.GeometricMultigridPoissonSolver(gpu_ctx, labels, w, levels, useGaussSeidel=True)
    x, it, hist = s.solveGeometricConjugateGradient(np.zeros_like(b), b, TOL, MAX_IT)
    s.close()
    assert abs(it - stored["iterations"]) <= 1
    m = min(len(hist), len(stored["history"]))
    assert m > 0 and max(abs(a - c) / a for a, c in zip(stored["history"][:m], hist[:m])) < 1e-5  # the north_star's bar


@pytest.mark.parametrize("dom,gs", [("simple", 0), ("complex", 0), ("complex", 1)])
def test_the_reference_smoother_test_against_the_restatement(testnode, port, dom, gs):
    """Test.cpp:1962-2105: 3 band sweeps, one interior sweep (damped Jacobi, or the four tiled Gauss-Seidel half-passes odd-forward,
    even-forward, even-backward, odd-backward), 3 band sweeps, then the residual's max (clamped at 0: the infNorm quirk) and L2 norm -- six
    rounds on the node's own domain and delta right-hand side, against the same sequence of restated operators."""
    n, rounds = 32, 6
    ok, log = testnode.run(gridSize=n, useComplexDomain=int(dom == "complex"), testSmoother=1, maxSmootherIterations=rounds, useGaussSeidelSmoothing=gs,
                           deltaFunctionAmplitude=AMPLITUDE)
    assert ok, log[-2000:]
    inf_node = [float(v) for v in re.findall(r"L-infinity norm: ([-+.\deE]+)", log)]
    l2_node = [float(v) for v in re.findall(r"L-2 norm: ([-+.\deE]+)", log)]
    assert len(inf_node) == rounds == len(l2_node)
    labels, w, levels, b = delta_problem(port.expand_domain, dom, n)
    cells = port.boundary_cells(labels, 3)
    x = np.zeros_like(b)
    for k in range(rounds):
        x = port.boundary_jacobi(x, b, labels, cells, 3, w)
        if gs:
            for odd, forward in ((True, True), (False, True), (False, False), (True, False)):
                x = port.gauss_seidel(x, b, labels, odd, forward, w)
        else:
            x = port.jacobi(x, b, labels, w)
        x = port.boundary_jacobi(x, b, labels, cells, 3, w)
        r = port.residual(x, b, labels, w)
        assert abs(port.inf_norm(r, labels) - inf_node[k]) <= 2e-9 * inf_node[k], k
        assert abs(np.sqrt(port.norm2(r, labels)) - l2_node[k]) <= 2e-9 * l2_node[k], k
    assert l2_node[-1] < l2_node[0]


@pytest.mark.parametrize("dom", ["simple", "complex"])
def test_the_reference_vcycle_convergence_test_against_the_restatement(testnode, port, dom):
    """Test.cpp:1877-1960: fifty damped-Jacobi V-cycles with `useInitialGuess` on a zero right-hand side, from the node's two-mode sinusoid
    (its cell positions are fpreal32: dx times the EXPANDED cell index), printing max(x, 0) and the L2 norm after each one -- the north_star's
    smoother mode in the reference's own convergence test."""
    n = 32
    ok, log = testnode.run(gridSize=n, useComplexDomain=int(dom == "complex"), testOneLevelVCycle=1)
    assert ok, log[-2000:]
    inf_node = [float(v) for v in re.findall(r"L-infinity norm: ([-+.\deE]+)", log)]
    l2_node = [float(v) for v in re.findall(r"L-2 norm: ([-+.\deE]+)", log)]
    assert len(inf_node) == 51 == len(l2_node)  # the initial guess, then fifty cycles
    bl, bw, dx = D.DOMAINS[dom](n)
    labels, w, off, levels = port.expand_domain(bl, bw)
    active = D.active_mask(labels)
    k, j, i = np.meshgrid(*[np.arange(s) for s in labels.shape], indexing="ij")
    px, py, pz = [(dx * a).astype(np.float32).astype(np.float64) for a in (i, j, k)]
    guess = np.sin(2 * np.pi * px) * np.sin(2 * np.pi * py) * np.sin(2 * np.pi * pz) + np.sin(4 * np.pi * px) * np.sin(4 * np.pi * py) * np.sin(4 * np.pi * py)
    x = np.where(active, guess, 0.0)
    s = port.solver(labels, w, levels, False)
    zero = np.zeros_like(x)
    for it in range(51):
        if it:
            x = s.vcycle(x, zero, True)
        assert abs(port.inf_norm(x, labels) - inf_node[it]) <= 5e-9 * inf_node[0], it
        assert abs(np.sqrt(port.norm2(x, labels)) - l2_node[it]) <= 5e-9 * l2_node[0], it
    assert l2_node[-1] < 1e-6 * l2_node[0]


@pytest.mark.parametrize("dom,n", [("simple", 16), ("complex", 24)])
def test_the_reference_diagonal_cg_test_against_the_restatement(testnode, port, dom, n):
    """The node's other CG branch (Test.cpp:838-1010, useMultigridPreconditioner off): its own diagonal preconditioner."""
    ok, log = testnode.run(gridSize=n, useComplexDomain=int(dom == "complex"), testConjugateGradient=1, useMultigridPreconditioner=0, solveCGGeometrically=1,
                           solverTolerance=TOL, maxSolverIterations=MAX_IT, deltaFunctionAmplitude=AMPLITUDE)
    assert ok, log[-2000:]
    hist_node = [float(v) for v in re.findall(r"Relative error: ([-+.\deE]+)", log)]
    it_node = int(re.findall(r"Iterations: (\d+)", log)[-1])
    labels, w, levels, b = delta_problem(port.expand_domain, dom, n)
    x, it, hist = port.solver(labels, w, levels, True).pcg(np.zeros_like(b), b, TOL, MAX_IT, diagonal=True)
    assert it == it_node and len(hist) == len(hist_node)
    assert max(abs(a - c) / c for a, c in zip(hist_node, hist)) < 1e-7  # a hundred CG steps amplify the last printed digit
