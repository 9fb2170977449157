"""The steps either side of the solve and the whole pressure projection against the reference's OWN node source --
HDK_GeometricFreeSurfacePressureSolver.cpp, compiled unmodified over oracle/shim/hdk_node_shim.h into oracle/_ref/libgmg_ref.so.

  CPU, where oracle/_ref exists: the C restatement (oracle/gmg_oracle.c) against the node's six builders one by one (GFS.cpp:746-1131, bit for
       bit) and against the whole solveGasSubclass (GFS.cpp:113-714, production wiring: tiled Gauss-Seidel V-cycle as the PCG preconditioner,
       or the diagonal one) -- iteration count, pressure, velocity.
  CPU, everywhere: the same against tests/golden/node_projection.npz, generated from the node by tests/golden/make_golden_frontend.py.
  GPU: the CUDA kernels (through the C ABI) against that fixture, with the bars of tests/test_frontend.py (labels and fpreal32 outputs bit for
       bit, fp64 outputs to 1e-14) and, for the whole projection, the north_star's (iteration count +-1, pressure to 1e-5)."""
import hashlib
import os
import re

import numpy as np
import pytest

from tests.common import GOLDEN_DIR
from tests.test_frontend import AIR, LIQUID, SOLID, make_fields, np_material_labels

CASES = [(24, 3), (32, 5), (20, 9)]
NODE_BUILDERS_CASE = (24, 3)
NODE_SOLVE_CASES = {"mg24": (24, 9, True), "mg32": (32, 4, True), "diag24": (24, 9, False)}
TOL, MAX_IT = 1e-7, 400


def nonfractional(cut):
    return [np.where(c > 0, 1.0, 0.0).astype(np.float32) for c in cut]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fixture():
    return np.load(os.path.join(GOLDEN_DIR, "node_projection.npz"))


# ---- one interface over the C restatement and the CUDA library ---------------------------------------------------------------------
class PortSteps:
    """The C restatement behind the method names of the CUDA library's Python mirror."""

    def __init__(self, port):
        self.p = port
        self.buildMaterialCellLabels = port.build_material_labels
        self.buildValidFaces = port.build_valid_faces
        self.buildMGDomainLabels = port.build_domain_labels
        self.buildMGBoundaryWeights = port.build_boundary_weights
        self.buildExpandedDomain = port.expand_domain
        self.buildRHS = port.build_rhs
        self.applyOldPressure = port.apply_old_pressure
        self.applySolutionToPressure = port.apply_solution_to_pressure
        self.applyPressureGradient = port.apply_pressure_gradient

    def solve(self, labels, w, levels, rhs, mg):
        s = self.p.solver(labels, w, levels, True)  # useGaussSeidel = true, GFS.cpp:463-466
        x, it, hist = s.pcg(np.zeros_like(rhs), rhs, TOL, MAX_IT, diagonal=not mg)
        return x, it


class GpuSteps:
    def __init__(self, ctx):
        self.ctx = ctx
        for name in ("buildMaterialCellLabels", "buildValidFaces", "buildMGDomainLabels", "buildMGBoundaryWeights", "buildExpandedDomain", "buildRHS",
                     "applyOldPressure", "applySolutionToPressure", "applyPressureGradient"):
            setattr(self, name, getattr(ctx, name))

    def solve(self, labels, w, levels, rhs, mg):
        from geometricmultigridpressuresolver_b200 import api

        s = api.GeometricMultigridPoissonSolver(self.ctx, labels, w, levels, useGaussSeidel=True)
        x, it, hist = s.solveGeometricConjugateGradient(np.zeros_like(rhs), rhs, TOL, MAX_IT, useMGPreconditioner=True if mg else "diagonal")
        s.close()
        return x, it


def check_builders_against_fixture(steps, exact_fp64):
    """The six builders on make_fields(24, 3) against the node's outputs.  exact_fp64: the restatement repeats the node's double operations
    in their order (bit for bit); the kernels may contract a multiply-add (1e-14, the bar of tests/test_frontend.py)."""
    g = fixture()
    n, seed = NODE_BUILDERS_CASE
    material, phi, cut, valid, vel, pressure = make_fields(n, seed)
    assert [sha(a) for a in [material, phi, pressure] + cut + valid + vel] == list(g["builders__inputs_sha"]), "the fixture's inputs no longer reproduce"

    def close(a, b):
        if exact_fp64:
            return (a == b).all()
        return np.abs(a - b).max() <= 1e-14 * max(np.abs(b).max(), 1e-300)

    bl = steps.buildMGDomainLabels(material)
    assert (bl == g["builders__domain_labels"]).all()
    bw = [steps.buildMGBoundaryWeights(cut[a], phi, valid[a], bl, a) for a in range(3)]
    for a in range(3):
        assert close(bw[a], g[f"builders__weights{a}"]), a
    labels, w, off, levels = steps.buildExpandedDomain(bl, [g[f"builders__weights{a}"] for a in range(3)])
    box = (slice(int(off[2]), int(off[2]) + n), slice(int(off[1]), int(off[1]) + n), slice(int(off[0]), int(off[0]) + n))
    outside = np.ones(labels.shape, dtype=bool)
    outside[box] = False
    rhs = steps.buildRHS(material, vel, cut, labels.shape, off)
    assert close(rhs[box], g["builders__rhs"]) and not rhs[outside].any()
    rhs_s = steps.buildRHS(material, vel, cut, labels.shape, off, [np.full_like(v, 0.25) for v in vel])
    assert close(rhs_s[box], g["builders__rhs_solid"])
    xo = steps.applyOldPressure(pressure, material, labels.shape, off)
    assert (xo[box] == g["builders__old_pressure"]).all() and not xo[outside].any()
    x = np.where(np.isin(labels, (0, 3)), np.random.default_rng(seed).random(labels.shape), 0.0)
    p = steps.applySolutionToPressure(np.full_like(pressure, -1.0), material, x, off)
    assert p.dtype == np.float32 and (p == g["builders__pressure"]).all()
    for a in range(3):
        v = steps.applyPressureGradient(vel[a], phi, pressure, valid[a], material, a)
        assert v.dtype == np.float32 and (v == g[f"builders__velocity{a}"]).all(), a


def check_projection_against_fixture(steps, name, p_tol, iteration_slack):
    """fields -> material labels -> valid faces -> domain labels / weights -> expanded domain -> rhs -> PCG -> pressure -> velocity against what
    the node's solveGasSubclass left behind on the same fields."""
    g = fixture()
    n, seed, mg = NODE_SOLVE_CASES[name]
    _, phi, cut, _, vel, _ = make_fields(n, seed)
    cut = nonfractional(cut)
    assert [sha(a) for a in [phi] + cut + vel] == list(g[f"{name}__inputs_sha"]), "the fixture's inputs no longer reproduce"
    dry = np.full(phi.shape, -10.0, dtype=np.float32)  # no solid field: the node substitutes a constant -10 dx (GFS.cpp:207-219)
    material = steps.buildMaterialCellLabels(phi, dry, cut)
    assert (material == np_material_labels(phi, dry, cut)).all() and {SOLID, LIQUID, AIR} >= set(np.unique(material))
    valid = [steps.buildValidFaces(material, cut[a], a) for a in range(3)]
    for a in range(3):
        assert (valid[a] == g[f"{name}__valid{a}"]).all(), a
    bl = steps.buildMGDomainLabels(material)
    bw = [steps.buildMGBoundaryWeights(cut[a], phi, valid[a], bl, a) for a in range(3)]
    labels, w, off, levels = steps.buildExpandedDomain(bl, bw)
    rhs = steps.buildRHS(material, vel, cut, labels.shape, off)
    x, it = steps.solve(labels, w, levels, rhs, mg)
    assert abs(it - int(g[f"{name}__iterations"])) <= iteration_slack, (it, int(g[f"{name}__iterations"]))
    p = steps.applySolutionToPressure(np.zeros(material.shape, np.float32), material, x, off)
    p_ref = g[f"{name}__pressure"]
    scale = np.abs(p_ref).max()
    assert scale > 0 and np.abs(p - p_ref).max() <= p_tol * scale
    for a in range(3):
        v_ref = g[f"{name}__velocity{a}"]
        v = steps.applyPressureGradient(vel[a], phi, p_ref, valid[a], material, a)
        assert (v == v_ref).all(), a  # from the node's own pressure: bit for bit
        v2 = steps.applyPressureGradient(vel[a], phi, p, valid[a], material, a)
        assert np.abs(v2 - v_ref).max() <= 2 * p_tol * max(np.abs(v_ref).max(), scale), a
    assert float(g[f"{name}__max_divergence"]) < 1e-4 * np.abs(rhs).max() + 1e-6  # the node's own check, GFS.cpp:662-707


# ---- CPU, against the node itself ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,seed", CASES)
def test_node_builders_against_the_restatement(port, ref, n, seed):
    material, phi, cut, valid, vel, pressure = make_fields(n, seed)
    bl = ref.node_domain_labels(material)
    assert (bl == port.build_domain_labels(material)).all()
    bw = []
    for axis in range(3):
        w_ref = ref.node_boundary_weights(cut[axis], phi, valid[axis], material, bl, axis)
        w_port = port.build_boundary_weights(cut[axis], phi, valid[axis], bl, axis)
        assert (w_ref == w_port).all(), axis  # same double operations in the same order: bit for bit
        bw.append(w_ref)
    labels, w, off, levels = ref.expand_domain(bl, bw)
    for sv in (None, [np.full_like(v, 0.25) for v in vel], [(np.arange(v.size, dtype=np.float32).reshape(v.shape) % 7) * np.float32(0.125) for v in vel]):
        r_ref = ref.node_rhs(material, vel, cut, labels, off, sv)
        r_port = port.build_rhs(material, vel, cut, labels.shape, off, sv)
        assert (r_ref == r_port).all()
        assert not r_ref[~np.isin(labels, (0, 3))].any()  # only active cells carry a right-hand side
    x_ref = ref.node_old_pressure(pressure, material, labels, off)
    assert (x_ref == port.apply_old_pressure(pressure, material, labels.shape, off)).all()
    x = np.where(np.isin(labels, (0, 3)), np.random.default_rng(seed).random(labels.shape), 0.0)
    p_ref = ref.node_solution_to_pressure(np.full_like(pressure, -1.0), material, x, labels, off)
    assert p_ref.dtype == np.float32 and (p_ref == port.apply_solution_to_pressure(np.full_like(pressure, -1.0), material, x, off)).all()
    for axis in range(3):
        v_ref = ref.node_pressure_gradient(vel[axis], cut[axis], phi, pressure, valid[axis], material, axis)
        v_port = port.apply_pressure_gradient(vel[axis], phi, pressure, valid[axis], material, axis)
        assert v_ref.dtype == np.float32 and (v_ref == v_port).all(), axis


@pytest.mark.parametrize("name", list(NODE_SOLVE_CASES))
def test_whole_node_against_the_chain_of_restated_pieces(port, ref, name):
    n, seed, mg = NODE_SOLVE_CASES[name]
    _, phi, cut, _, vel, _ = make_fields(n, seed)
    cut = nonfractional(cut)
    ok, p_ref, v_ref, valid_ref, log = ref.node_solve(phi, vel, cut, tolerance=TOL, max_iterations=MAX_IT, use_mg_preconditioner=mg)
    assert ok, log[-2000:]
    g = fixture()  # the committed fixture is this very run
    assert int(re.findall(r"Iterations: (\d+)", log)[-1]) == int(g[f"{name}__iterations"])
    assert (p_ref == g[f"{name}__pressure"]).all()
    for a in range(3):
        assert (v_ref[a] == g[f"{name}__velocity{a}"]).all() and (valid_ref[a] == g[f"{name}__valid{a}"]).all()


def test_node_with_a_solid_field_and_a_warm_start(port, ref):
    """The node's other inputs: a solid SDF with both signs (the second branch of isCellLiquid), a moving solid, useOldPressure."""
    n, seed = 24, 9
    _, phi, cut, _, vel, _ = make_fields(n, seed)
    cut = nonfractional(cut)
    solid = np.random.default_rng(1).random(phi.shape).astype(np.float32) - np.float32(0.5)
    sv = [np.full_like(v, 0.125) for v in vel]
    ok, p0, v0, valid0, log0 = ref.node_solve(phi, vel, cut, solid_surface=solid, solid_velocity=sv, tolerance=TOL, max_iterations=MAX_IT)
    assert ok, log0[-2000:]
    material = port.build_material_labels(phi, solid, cut)
    valid = [port.build_valid_faces(material, cut[a], a) for a in range(3)]
    for a in range(3):
        assert (valid[a] == valid0[a]).all()
    bl = port.build_domain_labels(material)
    bw = [port.build_boundary_weights(cut[a], phi, valid[a], bl, a) for a in range(3)]
    labels, w, off, levels = port.expand_domain(bl, bw)
    rhs = port.build_rhs(material, vel, cut, labels.shape, off, sv)
    s = port.solver(labels, w, levels, True)
    x, it, hist = s.pcg(np.zeros_like(rhs), rhs, TOL, MAX_IT)
    assert it == int(re.findall(r"Iterations: (\d+)", log0)[-1])
    p = port.apply_solution_to_pressure(np.zeros(material.shape, np.float32), material, x, off)
    assert np.abs(p - p0).max() <= 1e-6 * np.abs(p0).max()
    # warm start from the converged pressure: "Residual already below error" or a handful of iterations, same pressure
    ok, p1, v1, valid1, log1 = ref.node_solve(phi, vel, cut, solid_surface=solid, solid_velocity=sv, pressure=p0, tolerance=1e-5, max_iterations=MAX_IT, use_old_pressure=True)
    assert ok and np.abs(p1 - p0).max() <= 1e-4 * np.abs(p0).max()
    x0 = port.apply_old_pressure(p0, material, labels.shape, off)
    xw, itw, histw = s.pcg(x0, rhs, 1e-5, MAX_IT)
    found = re.findall(r"Iterations: (\d+)", log1)
    assert (itw == -1 and not found) or (found and itw == int(found[-1]))


# ---- CPU, against the committed fixture ---------------------------------------------------------------------------------------------
def test_restated_builders_against_the_node_fixture(port):
    check_builders_against_fixture(PortSteps(port), exact_fp64=True)


@pytest.mark.parametrize("name", list(NODE_SOLVE_CASES))
def test_restated_projection_against_the_node_fixture(port, name):
    check_projection_against_fixture(PortSteps(port), name, p_tol=1e-6, iteration_slack=0)


# ---- GPU, against the committed fixture ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_frontend_kernels_against_the_node_fixture(gpu_ctx):
    check_builders_against_fixture(GpuSteps(gpu_ctx), exact_fp64=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mg24", "mg32"])
def test_projection_against_the_node_fixture(gpu_ctx, name):
    """The node's production wiring on the GPU: tiled Gauss-Seidel V-cycle inside the PCG, every step either side of it a CUDA kernel."""
    check_projection_against_fixture(GpuSteps(gpu_ctx), name, p_tol=1e-5, iteration_slack=1)


def test_no_reference_assert_trips_on_the_fixture_fields():
    """The same calls through oracle/_ref/libgmg_ref_dbg.so -- the reference's sources with their asserts LIVE (no NDEBUG): every precondition the
    node checks on its own data (weights > 0 on valid faces, liquid / air pairing at the free surface, active labels under every liquid cell,
    matching resolutions, ...) holds on the seeded fields these tests use."""
    from oracle import bindings

    if not os.path.exists(bindings.REF_DBG_SO):
        pytest.skip("oracle/_ref/libgmg_ref_dbg.so not present (built only where /root/reference exists)")
    dbg = bindings.RefLib(debug=True)
    n, seed = NODE_BUILDERS_CASE
    material, phi, cut, valid, vel, pressure = make_fields(n, seed)
    bl = dbg.node_domain_labels(material)
    bw = [dbg.node_boundary_weights(cut[a], phi, valid[a], material, bl, a) for a in range(3)]
    labels, w, off, levels = dbg.expand_domain(bl, bw)
    dbg.node_rhs(material, vel, cut, labels, off, [np.full_like(v, 0.25) for v in vel])
    x = dbg.node_old_pressure(pressure, material, labels, off)
    dbg.node_solution_to_pressure(pressure, material, x, labels, off)
    for a in range(3):
        dbg.node_pressure_gradient(vel[a], cut[a], phi, pressure, valid[a], material, a)
        dbg.build_valid_faces(material, cut[a], a)
    assert (dbg.build_material_labels(phi, np.full(phi.shape, -1.0, np.float32), cut) == material).all()
    _, phi, cut, _, vel, _ = make_fields(24, 9)
    ok, p, v, vf, log = dbg.node_solve(phi, vel, nonfractional(cut), tolerance=TOL, max_iterations=MAX_IT)
    assert ok and (p == fixture()["mg24__pressure"]).all()  # -O1 with asserts, -O3 without: the same pressure, bit for bit
