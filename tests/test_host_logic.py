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


def _plane_extents(active):
    """(planes, 4) bounding rectangles (x0, x1, y0, y1) of a (z, y, x) boolean mask, as k_plane_extents produces them."""
    ext = np.zeros((active.shape[0], 4), dtype=np.int32)
    for z in range(active.shape[0]):
        ys, xs = np.nonzero(active[z])
        ext[z] = (xs.min(), xs.max() + 1, ys.min(), ys.max() + 1) if len(xs) else (active.shape[2], 0, active.shape[1], 0)
    return ext


@pytest.mark.parametrize("name", ["flipsplash", "sphere", "narrow_band", "liquid_box"])
def test_transfer_plan_covers_every_active_cell_and_little_else(name):
    """gmg_transfer_plan (pure host): the groups tile the non-empty planes in order, every active cell lies inside its plane's
    group, and merging never inflates a group beyond 15 % of its members' own rectangles."""
    labels, w, dx = D.DOMAINS[name](32)
    active = labels == D.INTERIOR
    ext = _plane_extents(active)
    groups, cells = api.transfer_plan(ext)
    assert cells == sum(int(g[1] - g[0]) * int(g[3] - g[2]) * int(g[5] - g[4]) for g in groups)
    assert active.sum() <= cells <= active.size
    covered = np.zeros_like(active)
    last = 0
    for z0, z1, x0, x1, y0, y1 in groups:
        assert z0 >= last and z1 > z0 and x1 > x0 and y1 > y0
        last = z1
        covered[z0:z1, y0:y1, x0:x1] = True
        own = sum(int(e[1] - e[0]) * int(e[3] - e[2]) for e in ext[z0:z1])
        assert all(e[1] > e[0] for e in ext[z0:z1])  # no empty plane inside a group
        assert (z1 - z0) * (x1 - x0) * (y1 - y0) <= 1.15 * own + 1e-9
    assert covered[active].all()
    empty = np.nonzero(ext[:, 1] <= ext[:, 0])[0]
    assert not covered[empty].any()


def test_transfer_plan_edge_cases():
    groups, cells = api.transfer_plan(np.zeros((0, 4), dtype=np.int32))
    assert len(groups) == 0 and cells == 0
    groups, cells = api.transfer_plan(np.array([[5, 0, 5, 0]] * 3, dtype=np.int32))  # all planes empty
    assert len(groups) == 0 and cells == 0
    ext = np.array([[2, 6, 1, 4], [2, 6, 1, 4], [9, 0, 9, 0], [0, 40, 0, 40], [10, 12, 10, 12]], dtype=np.int32)
    groups, cells = api.transfer_plan(ext)
    # identical planes merge; the empty plane splits; a tiny rectangle is not merged into a huge one (it would waste > 15 %)
    assert groups.tolist() == [[0, 2, 2, 6, 1, 4], [3, 4, 0, 40, 0, 40], [4, 5, 10, 12, 10, 12]]
    assert cells == 2 * 12 + 1600 + 4


def test_gather_face_weights_matches_direct_indexing():
    """gmg_gather_face_weights (pure host, what the constructor runs beside the hierarchy build): the six face weights of a cell
    list, read where the caller's expanded weight grids lie, 0 outside the hint box (+1 on a face's own axis)."""
    rng = np.random.default_rng(3)
    res = (20, 18, 16)  # x, y, z
    w = [rng.random((res[2] + (a == 2), res[1] + (a == 1), res[0] + (a == 0))) for a in range(3)]
    org, n = (4, 2, 6), (12, 10, 8)  # a storage box inside the grid
    pitch, plane = 16, 16 * n[1]
    cells = np.array([(x, y, z) for z in range(1, n[2] - 1) for y in range(1, n[1] - 1) for x in range(1, n[0] - 1)])
    cells = cells[rng.permutation(len(cells))[:300]]
    idx = cells[:, 2] * plane + cells[:, 1] * pitch + cells[:, 0]

    def expect(bounds):
        out = np.zeros((6, len(cells)))
        for k, (x, y, z) in enumerate(cells):
            e = np.array([x + org[0], y + org[1], z + org[2]])
            for f in range(6):
                a = f >> 1
                p = e.copy()
                p[a] += f & 1
                lo = np.array(bounds[:3]) if bounds is not None else np.zeros(3, dtype=int)
                hi = np.array(bounds[3:]) if bounds is not None else np.array(res)
                hi = hi.copy()
                hi[a] += 1
                if (p >= lo).all() and (p < hi).all():
                    out[f, k] = w[a][p[2], p[1], p[0]]
        return out

    got = api.gather_face_weights(idx, pitch, plane, org, w)
    assert np.array_equal(got, expect(None))
    bounds = (6, 4, 8, 13, 10, 12)  # a hint box cutting through the cell list
    got = api.gather_face_weights(idx, pitch, plane, org, w, bounds)
    want = expect(bounds)
    assert np.array_equal(got, want) and (want == 0).any() and (want != 0).any()


def test_plane_wise_narrow_band_labels_match_the_dense_pipeline(port):
    """bench.py's narrow1024 block builds the one-byte expanded labels plane by plane (no dense float grid); at a size the
    oracle handles it must equal buildExpandedCellLabels + setBoundaryCellLabels on narrow_band_domain's arrays."""
    import bench

    for n, t in [(32, 4), (48, 12)]:
        bl, bw, dx = D.narrow_band_domain(n, t)
        labels, w, off, levels = port.expand_domain(bl, bw)
        l8, off8, lev8, box = bench.narrow_band_labels_u8(n, t)
        sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
        assert lev8 == levels and list(off8) == list(off)
        assert (labels[sl] == l8[sl]).all()
        outside = np.ones(labels.shape, dtype=bool)
        outside[sl] = False
        assert (labels[outside] == D.EXTERIOR).all()
