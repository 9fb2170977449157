"""CPU: plain-C oracle vs the reference's own sources (oracle/_ref, built where /root/reference exists).

This is what pins the oracle: every operator, both smoothers, the builders, the V-cycle and PCG are run
on seeded inputs through BOTH libraries and must agree to fp64 round-off / bit-exactly.
Skipped where oracle/_ref is absent (the committed fixtures in tests/golden then carry the pin)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import domains as D
from tests.common import relerr

TOL = 1e-12
DOMAINS = [("sphere", 20, {}), ("complex", 24, {}), ("flipsplash", 32, {"shape": (32, 20, 28)}), ("narrow_band", 32, {"thickness": 5})]


@pytest.mark.parametrize("dom,n,kw", DOMAINS)
def test_builders_bit_exact(port, ref, dom, n, kw):
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    l1, w1, o1, lv1 = ref.expand_domain(bl, bw)
    l2, w2, o2, lv2 = port.expand_domain(bl, bw)
    assert lv1 == lv2 and (o1 == o2).all() and l1.shape == l2.shape
    assert (l1 == l2).all()
    for a in range(3):
        assert (w1[a] == w2[a]).all()
    assert ref.unit_test_boundary_cells(l1, w1) and ref.unit_test_exterior_cells(l1)
    lab = l1
    for level in range(lv1 - 1):
        c1, c2 = ref.coarsen_labels(lab), port.coarsen_labels(lab)
        assert (c1 == c2).all()
        assert ref.unit_test_coarsening(c1, lab) == port.unit_test_coarsening(c1, lab)
        for width in (1, 2, 3):
            b1, b2 = ref.boundary_cells(c1, width), port.boundary_cells(c1, width)
            assert b1.shape == b2.shape and (b1 == b2).all()
        lab = c1
        if not D.active_mask(lab).any():
            break
    b1, b2 = ref.boundary_cells(l1, 3), port.boundary_cells(l1, 3)
    assert b1.shape == b2.shape and (b1 == b2).all()


@pytest.mark.parametrize("dom,n,kw", DOMAINS)
def test_operators(port, ref, dom, n, kw):
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, lv = ref.expand_domain(bl, bw)
    x, b = D.random_active(labels, 1), D.random_active(labels, 2)
    cells = ref.boundary_cells(labels, 3)
    for ww in (w, None):
        assert relerr(port.jacobi(x, b, labels, ww), ref.jacobi(x, b, labels, ww)) < TOL
        assert relerr(port.boundary_jacobi(x, b, labels, cells, 2, ww), ref.boundary_jacobi(x, b, labels, cells, 2, ww)) < TOL
        assert relerr(port.apply(x, labels, ww), ref.apply(x, labels, ww)) < TOL
        assert relerr(port.residual(x, b, labels, ww), ref.residual(x, b, labels, ww)) < TOL
        for odd in (0, 1):
            for fwd in (0, 1):
                assert relerr(port.gauss_seidel(x, b, labels, odd, fwd, ww), ref.gauss_seidel(x, b, labels, odd, fwd, ww)) < TOL
    cl = ref.coarsen_labels(labels)
    assert relerr(port.downsample(x, cl, labels), ref.downsample(x, cl, labels)) < TOL
    xc = D.random_active(cl, 3)
    assert relerr(port.upsample_add(x, xc, labels, cl), ref.upsample_add(x, xc, labels, cl)) < TOL
    assert abs(port.dot(x, b, labels) - ref.dot(x, b, labels)) < 1e-12 * ref.norm2(x, labels)
    assert abs(port.norm2(x, labels) - ref.norm2(x, labels)) < 1e-13 * ref.norm2(x, labels)
    assert port.inf_norm(x, labels) == ref.inf_norm(x, labels)
    assert relerr(port.axpy(x, b, 0.37, labels), ref.axpy(x, b, 0.37, labels)) < TOL
    assert relerr(port.add_scaled(x, b, -1.0, labels), ref.add_scaled(x, b, -1.0, labels)) < TOL
    assert relerr(port.scale(x, 2.5, labels), ref.scale(x, 2.5, labels)) < TOL


@pytest.mark.parametrize("use_gs", [False, True])
@pytest.mark.parametrize("dom,n,kw", DOMAINS[:3])
def test_vcycle_and_pcg(port, ref, dom, n, kw, use_gs):
    bl, bw, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, lv = ref.expand_domain(bl, bw)
    s1, s2 = ref.solver(labels, w, lv, use_gs), port.solver(labels, w, lv, use_gs)
    assert s1.levels == s2.levels
    b = D.random_rhs(labels, dx, 5)
    assert relerr(s2.vcycle(np.zeros_like(b), b), s1.vcycle(np.zeros_like(b), b)) < 1e-11
    x1, it1, h1 = s1.pcg(np.zeros_like(b), b, 1e-6, 200)
    x2, it2, h2 = s2.pcg(np.zeros_like(b), b, 1e-6, 200)
    assert it1 == it2 and len(h1) == len(h2)
    assert (np.abs(h1 - h2) / h1).max() < 1e-8
    assert relerr(x2, x1) < 1e-10


def test_coarse_matrix_job_count_quirk(port):
    """With J UT_ThreadedAlgorithm jobs the reference assembles J*A at the coarsest level (MG.cpp:334-389 does not
    split the tile range); the oracle's coarse_scale = J reproduces it.  The shim reads GMG_SHIM_JOBS at load, so
    the J=3 reference run happens in a child process."""
    from oracle import bindings

    if not os.path.exists(bindings.REF_SO):
        pytest.skip("oracle/_ref not built")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle.bindings import RefLib\n"
        "from geometricmultigridpressuresolver_b200 import domains as D\n"
        "ref = RefLib(); bl, bw, dx = D.sphere_domain(20)\n"
        "labels, w, off, lv = ref.expand_domain(bl, bw)\n"
        "b = D.random_rhs(labels, dx, 5); s = ref.solver(labels, w, lv, False, coarse_scale=3.0)\n"
        "np.save(sys.argv[1], s.vcycle(np.zeros_like(b), b))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "v.npy")
        env = dict(os.environ, GMG_SHIM_JOBS="3")
        subprocess.check_call([sys.executable, "-c", code, out], env=env)
        v_ref = np.load(out)
    bl, bw, dx = D.sphere_domain(20)
    labels, w, off, lv = port.expand_domain(bl, bw)
    b = D.random_rhs(labels, dx, 5)
    v3 = port.solver(labels, w, lv, False, coarse_scale=3.0).vcycle(np.zeros_like(b), b)
    v1 = port.solver(labels, w, lv, False, coarse_scale=1.0).vcycle(np.zeros_like(b), b)
    assert relerr(v3, v_ref) < 1e-11
    assert relerr(v1, v_ref) > 1e-3  # and it really is a different operator


@pytest.mark.parametrize("shape,seed", [((24, 24, 24), 3), ((40, 18, 33), 5), ((16, 16, 16), 1), ((1, 5, 3), 2), ((17, 1, 35), 4), ((2, 2, 2), 6)])
def test_material_labels_and_valid_faces_against_the_reference_sources(port, ref, shape, seed):
    """buildMaterialCellLabels is the reference's own HDK_Utilities.cpp (compiled unmodified over the shim's SIM_RawField), buildValidFaces its
    own templates findOccupiedFaceTiles / uncompressTiles / classifyValidFaces in the order of GFS.cpp:717-744: this pins the C restatement the
    GPU kernels are held to (tests/test_frontend.py).  Shapes span several 16^3 tiles, partial tiles and one-cell-thick axes."""
    rng = np.random.default_rng(seed)
    phi = (rng.random(shape).astype(np.float32) - 0.45)
    solid = (rng.random(shape).astype(np.float32) - 0.5)
    cut = []
    for a in range(3):
        fs = D.face_shape(shape, a)
        w = (rng.random(fs) < 0.65).astype(np.float32) * (rng.random(fs).astype(np.float32) * 0.95 + 0.05)
        cut.append(w)
    # a solid block with every face closed, so whole tiles of the label field stay constant SOLID
    blk = tuple(slice(0, max(1, s // 2)) for s in shape)
    for a in range(3):
        na = 2 - a
        sl = list(blk)
        sl[na] = slice(0, blk[na].stop + 1)
        cut[a][tuple(sl)] = 0.0
    for so in (solid, np.full(shape, -1.0, np.float32), np.zeros(shape, np.float32)):
        m_ref = ref.build_material_labels(phi, so, cut)
        m_port = port.build_material_labels(phi, so, cut)
        assert (m_ref == m_port).all()
        assert (m_ref[blk] == 0).all()
        for axis in range(3):
            v_ref = ref.build_valid_faces(m_ref, cut[axis], axis)
            v_port = port.build_valid_faces(m_ref, cut[axis], axis)
            assert v_ref.dtype == np.float32 and (v_ref == v_port).all()
    assert set(np.unique(m_ref)) <= {0, 1, 2}


def test_material_labels_of_the_frontend_test_fields_against_the_reference_sources(port, ref):
    from tests.test_frontend import make_fields, make_solid_sdf

    for n, seed in ((24, 3), (32, 5), (32, 9)):
        material, phi, cut, valid, vel, pressure = make_fields(n, seed)
        dry = np.full(phi.shape, -1.0, dtype=np.float32)
        assert (ref.build_material_labels(phi, dry, cut) == material).all()
        wet = make_solid_sdf(phi.shape)
        assert (ref.build_material_labels(phi, wet, cut) == port.build_material_labels(phi, wet, cut)).all()
        for axis in range(3):
            assert (ref.build_valid_faces(material, cut[axis], axis) == valid[axis]).all()
