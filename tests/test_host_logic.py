"""CPU: host-side logic that needs no device -- the expansion formula, domain generators, array conventions."""
import ctypes as C

import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import api
from geometricmultigridpressuresolver_b200 import domains as D

# SURVEY.md appendix D (from Ops.h:1340-1360)
TABLE_D = [((64, 64, 64), 5, 16, (128, 128, 128)), ((128, 128, 128), 6, 32, (256, 256, 256)), ((256, 256, 256), 7, 64, (512, 512, 512)),
           ((300, 200, 300), 7, 64, (512, 512, 512)), ((512, 512, 512), 8, 128, (1024, 1024, 1024)), ((1024, 1024, 1024), 9, 256, (2048, 2048, 2048)),
           ((24, 24, 24), 4, 8, (64, 64, 64)), ((20, 32, 28), 4, 8, (64, 64, 64)), ((16, 16, 16), 3, 4, (32, 32, 32))]


@pytest.mark.parametrize("base,levels,pad,exp", TABLE_D)
def test_expand_dims_table(port, base, levels, pad, exp):
    lib = api.load_library()
    b, e, o, lv = (C.c_int64 * 3)(*base), (C.c_int64 * 3)(), (C.c_int64 * 3)(), C.c_int()
    assert lib.gmg_expand_dims(b, e, o, C.byref(lv)) == 0
    assert (tuple(e), tuple(o), lv.value) == (exp, (pad,) * 3, levels)
    shape, off, plv = port.expand_dims(base[::-1])
    assert (shape[::-1], tuple(off), plv) == (exp, (pad,) * 3, levels)
    shape2, off2, lv2 = api.expand_dims(base[::-1])  # the Python mirror's host-only helper (used by buildExpandedDomainLazy)
    assert (shape2[::-1], tuple(off2), lv2) == (exp, (pad,) * 3, levels)


def test_ghost_fluid_theta_matches_reference_cases():
    # HDK_Utilities.h:25-42
    phi0 = np.array([-1.0, -1.0, 1.0, 1.0, -0.25])
    phi1 = np.array([-2.0, 3.0, -1.0, 2.0, 0.0])
    np.testing.assert_allclose(D.ghost_fluid_theta(phi0, phi1), [1.0, 0.25, 0.5, 0.0, 1.0])


@pytest.mark.parametrize("name", sorted(D.DOMAINS))
def test_domains_are_well_formed(name, port):
    n = 24
    labels, w, dx = D.DOMAINS[name](n)
    assert labels.dtype == np.int32 and labels.shape == (n, n, n)
    assert set(np.unique(labels)) <= {D.INTERIOR, D.EXTERIOR, D.DIRICHLET}
    for a in range(3):
        assert w[a].shape == D.face_shape(labels.shape, a) and (w[a] >= 0).all() and np.isfinite(w[a]).all()
        assert w[a].max() <= 100.0  # theta clamp .01
    exp, we, off, lv = port.expand_domain(labels, w)
    # the reference's own invariant checkers accept what the generators make
    assert port.unit_test_boundary_cells(exp, we) and port.unit_test_exterior_cells(exp)
    # faces of EXTERIOR cells carry no weight (asserted by the reference at Ops.h:253-254)
    ext = exp == D.EXTERIOR
    assert not we[0][:, :, :-1][ext].any() and not we[0][:, :, 1:][ext].any()
    assert not we[1][:, :-1, :][ext].any() and not we[1][:, 1:, :][ext].any()
    assert not we[2][:-1, :, :][ext].any() and not we[2][1:, :, :][ext].any()


def test_rhs_builders_respect_the_zero_invariant(port):
    labels, w, dx = D.sphere_domain(24)
    exp, we, off, lv = port.expand_domain(labels, w)
    for b in (D.random_rhs(exp, dx), D.delta_rhs(exp, [int(off[0]) + 12] * 3, dx), D.random_active(exp, 3)):
        assert b.any() and not b[~D.active_mask(exp)].any()
