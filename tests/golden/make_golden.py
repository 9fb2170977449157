"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

The fixtures are outputs of the reference's own hot-path sources (compiled unmodified into
oracle/_ref/libgmg_ref.so against oracle/shim -- the HDK and Eigen are not installable offline) on
the synthetic inputs of geometricmultigridpressuresolver_b200/domains.py.  /root/reference cannot
travel to the GPU box, these small files can.  Re-run:  python tests/golden/make_golden.py

Each case stores: expanded labels, per-level labels and boundary-cell lists (bit-exact targets),
the Jacobi-mode PCG residual history / iteration count / solution (cropped to the base box), one
V-cycle on a seeded random rhs, and one application of every grid operator on seeded inputs.
The debug build (asserts live) is run once per case to prove no reference precondition trips.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402
from oracle.bindings import RefLib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (domain, size argument, kwargs, store full grids?)
CASES = {
    "simple16": ("simple", 16, {}, True),
    "complex16": ("complex", 16, {}, True),
    "sphere24": ("sphere", 24, {}, True),          # non power of two: exercises the padding formula
    "flipsplash_24x16x24": ("flipsplash", 24, {"shape": (24, 16, 24)}, True),  # non-cubic
    "liquid_box16": ("liquid_box", 16, {}, True),
    "narrow_band32": ("narrow_band", 32, {"thickness": 4}, True),
    "sphere32": ("sphere", 32, {}, True),          # level cap (MG.cpp:243-248) triggers here
    "sphere64": ("sphere", 64, {}, False),         # BASELINE config 1
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def crop(a, off, base_shape):
    return a[off[2] : off[2] + base_shape[0], off[1] : off[1] + base_shape[1], off[0] : off[0] + base_shape[2]]


def rhs_for(labels, off, base_shape, dx):
    c = [int(off[0]) + base_shape[2] // 2, int(off[1]) + base_shape[1] // 2, int(off[2]) + base_shape[0] // 2]
    b = D.delta_rhs(labels, c, dx)
    if not b.any():
        b = D.random_rhs(labels, dx)
    return b


def build_case(ref: RefLib, dbg: RefLib, name: str):
    dom, n, kw, full = CASES[name]
    base_labels, base_w, dx = D.DOMAINS[dom](n, **kw)
    labels, w, off, levels = ref.expand_domain(base_labels, base_w)
    out = {
        "dx": dx, "offset": off, "mg_levels": levels, "base_shape": np.array(base_labels.shape),
        "base_labels_sha": sha(base_labels.astype(np.int32)), "base_w_sha": "".join(sha(a) for a in base_w),
        "labels_sha": sha(labels.astype(np.int32)), "weights_sha": "".join(sha(a) for a in w),
    }
    assert ref.unit_test_boundary_cells(labels, w) and ref.unit_test_exterior_cells(labels)
    # the asserts-on build must walk the same path without tripping
    dl, dw, doff, dlev = dbg.expand_domain(base_labels, base_w)
    assert (dl == labels).all() and dlev == levels
    solver = ref.solver(labels, w, levels, use_gs=False)
    dsolver = dbg.solver(labels, w, levels, use_gs=False)
    nl = solver.levels
    assert dsolver.levels == nl
    out["solver_levels"] = nl
    lv_labels = [labels]
    for l in range(1, nl):
        lv_labels.append(ref.coarsen_labels(lv_labels[-1]))
        assert ref.unit_test_coarsening(lv_labels[l], lv_labels[l - 1])
    for l in range(nl):
        cells = ref.boundary_cells(lv_labels[l], 3)
        out[f"labels_sha_L{l}"] = sha(lv_labels[l].astype(np.int32))
        out[f"cells_sha_L{l}"] = sha(cells.astype(np.int64))
        out[f"cells_count_L{l}"] = len(cells)
        if full:
            out[f"labels_L{l}"] = lv_labels[l].astype(np.int8)
            out[f"cells_L{l}"] = cells.astype(np.int16)
    if full:
        out["labels"] = labels.astype(np.int8)
    # PCG, Jacobi smoother (north_star), zero guess, tol 1e-6
    b = rhs_for(labels, off, base_labels.shape, dx)
    x, iters, hist = solver.pcg(np.zeros_like(b), b, 1e-6, 1000)
    dx_, diters, dhist = dsolver.pcg(np.zeros_like(b), b, 1e-6, 1000)
    assert diters == iters and np.allclose(dhist, hist, rtol=1e-9)
    out["pcg_iterations"] = iters
    out["pcg_history"] = hist
    out["rhs_sha"] = sha(b)
    xc = crop(x, off, base_labels.shape)
    out["pcg_x"] = xc if full else xc[::4, ::4, ::4].copy()
    out["pcg_x_norm2"] = float((x * x).sum())
    # the node's other mode: diagonal-preconditioned CG (GFS.cpp:485-618), capped at 60 iterations (slow convergence)
    xd, ditd, dhd = solver.pcg(np.zeros_like(b), b, 1e-6, 60, diagonal=True)
    out["dpcg_iterations"] = ditd
    out["dpcg_history"] = dhd
    xdc = crop(xd, off, base_labels.shape)
    out["dpcg_x"] = xdc if full else xdc[::4, ::4, ::4].copy()
    # one V-cycle on a seeded random rhs, and with an initial guess
    rb = D.random_rhs(labels, dx, seed=7)
    v = solver.vcycle(np.zeros_like(rb), rb)
    vc = crop(v, off, base_labels.shape)
    out["vcycle_x"] = vc if full else vc[::4, ::4, ::4].copy()
    out["vcycle_norm2"] = float((v * v).sum())
    if full:
        x0 = D.random_active(labels, 11, scale=dx * dx)
        out["vcycle_guess_x"] = crop(solver.vcycle(x0, rb, use_initial_guess=True), off, base_labels.shape)
        # every grid operator once, seeded inputs
        xs = D.random_active(labels, 1)
        bs = D.random_active(labels, 2)
        cells0 = ref.boundary_cells(labels, 3)
        out["op_jacobi"] = crop(ref.jacobi(xs, bs, labels, w), off, base_labels.shape)
        out["op_band3"] = crop(ref.boundary_jacobi(xs, bs, labels, cells0, 3, w), off, base_labels.shape)
        out["op_apply"] = crop(ref.apply(xs, labels, w), off, base_labels.shape)
        out["op_residual"] = crop(ref.residual(xs, bs, labels, w), off, base_labels.shape)
        out["op_gs_odd_fwd"] = crop(ref.gauss_seidel(xs, bs, labels, 1, 1, w), off, base_labels.shape)
        out["op_gs_even_bwd"] = crop(ref.gauss_seidel(xs, bs, labels, 0, 0, w), off, base_labels.shape)
        out["op_dot"] = ref.dot(xs, bs, labels)
        out["op_norm2"] = ref.norm2(xs, labels)
        out["op_inf_norm"] = ref.inf_norm(xs, labels)
        if nl > 1:
            l1 = lv_labels[1]
            out["op_downsample"] = ref.downsample(xs, l1, labels)
            xc1 = D.random_active(l1, 3)
            out["op_upsample"] = crop(ref.upsample_add(xs, xc1, labels, l1), off, base_labels.shape)
            b1 = D.random_active(l1, 4)
            out["op_jacobi_L1"] = ref.jacobi(xc1, b1, l1)
            out["op_band3_L1"] = ref.boundary_jacobi(xc1, b1, l1, ref.boundary_cells(l1, 3), 3)
    solver.close()
    dsolver.close()
    return out


def main():
    ref, dbg = RefLib(), RefLib(debug=True)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        out = build_case(ref, dbg, name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: levels {out['mg_levels']} -> {out['solver_levels']}, pcg iterations {out['pcg_iterations']}, "
              f"final rel res {out['pcg_history'][-1]:.3e}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
