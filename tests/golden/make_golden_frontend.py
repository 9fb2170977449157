"""Generate tests/golden/frontend_fields.npz from the REFERENCE ITSELF (run in the build container only).

buildMaterialCellLabels is the reference's own HDK_Utilities.cpp, compiled unmodified into oracle/_ref/libgmg_ref.so over the shim's
SIM_RawField / SIM_RawIndexField; the valid-face flags come from its own templates findOccupiedFaceTiles / uncompressTiles /
classifyValidFaces (HDK_Utilities.h) called in the order of HDK_GeometricFreeSurfacePressureSolver.cpp:717-744.  Inputs are the seeded
fields of tests/common.py:frontend_fields (their sha256 is stored, so a drifting generator is noticed).  Re-run:
python tests/golden/make_golden_frontend.py
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import RefLib  # noqa: E402
from tests.common import FRONTEND_CASES, frontend_fields  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = RefLib()
    out = {}
    for name in FRONTEND_CASES:
        phi, solid, cut = frontend_fields(name)
        out[f"{name}__inputs_sha"] = np.array([sha(phi), sha(solid)] + [sha(c) for c in cut])
        material = ref.build_material_labels(phi, solid, cut)
        out[f"{name}__material"] = material.astype(np.int8)
        for axis in range(3):
            v = ref.build_valid_faces(material, cut[axis], axis)
            assert set(np.unique(v)) <= {0.0, 1.0}
            out[f"{name}__valid{axis}"] = v.astype(np.uint8)
        print(name, material.shape, "SOLID/LIQUID/AIR", [(material == k).sum() for k in range(3)], "valid faces", [int(out[f"{name}__valid{a}"].sum()) for a in range(3)])
    np.savez_compressed(os.path.join(HERE, "frontend_fields.npz"), **out)


def node_fixture():
    """tests/golden/node_projection.npz: outputs of the reference's own NODE source (HDK_GeometricFreeSurfacePressureSolver.cpp, compiled
    unmodified over oracle/shim/hdk_node_shim.h) on the seeded fields of tests/test_frontend.py:make_fields -- its six builders one by one
    and the whole solveGasSubclass (tiled Gauss-Seidel V-cycle preconditioner, and the diagonal one)."""
    from tests.test_frontend import make_fields
    from tests.test_node_reference import NODE_BUILDERS_CASE, NODE_SOLVE_CASES, TOL, MAX_IT, nonfractional

    ref = RefLib()
    out = {}
    n, seed = NODE_BUILDERS_CASE
    material, phi, cut, valid, vel, pressure = make_fields(n, seed)
    out["builders__inputs_sha"] = np.array([sha(a) for a in [material, phi, pressure] + cut + valid + vel])
    bl = ref.node_domain_labels(material)
    bw = [ref.node_boundary_weights(cut[a], phi, valid[a], material, bl, a) for a in range(3)]
    labels, w, off, levels = ref.expand_domain(bl, bw)
    sv = [np.full_like(v, 0.25) for v in vel]
    x = np.where(np.isin(labels, (0, 3)), np.random.default_rng(seed).random(labels.shape), 0.0)
    box = (slice(int(off[2]), int(off[2]) + n), slice(int(off[1]), int(off[1]) + n), slice(int(off[0]), int(off[0]) + n))
    out["builders__domain_labels"] = bl.astype(np.int8)
    for a in range(3):
        out[f"builders__weights{a}"] = bw[a]
        out[f"builders__velocity{a}"] = ref.node_pressure_gradient(vel[a], cut[a], phi, pressure, valid[a], material, a)
    out["builders__rhs"] = ref.node_rhs(material, vel, cut, labels, off)[box]
    out["builders__rhs_solid"] = ref.node_rhs(material, vel, cut, labels, off, sv)[box]
    out["builders__old_pressure"] = ref.node_old_pressure(pressure, material, labels, off)[box]
    out["builders__pressure"] = ref.node_solution_to_pressure(np.full_like(pressure, -1.0), material, x, labels, off)
    for name, (n, seed, mg) in NODE_SOLVE_CASES.items():
        _, phi, cut, _, vel, _ = make_fields(n, seed)
        cut = nonfractional(cut)
        ok, p, v, valid, log = ref.node_solve(phi, vel, cut, tolerance=TOL, max_iterations=MAX_IT, use_mg_preconditioner=mg)
        assert ok
        import re

        out[f"{name}__inputs_sha"] = np.array([sha(a) for a in [phi] + cut + vel])
        out[f"{name}__iterations"] = np.array(int(re.findall(r"Iterations: (\d+)", log)[-1]))
        out[f"{name}__max_divergence"] = np.array(float(re.search(r"Max divergence: ([-+.\deE]+)", log).group(1)))
        out[f"{name}__pressure"] = p
        for a in range(3):
            out[f"{name}__velocity{a}"] = v[a]
            out[f"{name}__valid{a}"] = valid[a].astype(np.uint8)
        print(name, "iterations", out[f"{name}__iterations"], "max divergence", out[f"{name}__max_divergence"], "max |p|", np.abs(p).max())
    np.savez_compressed(os.path.join(HERE, "node_projection.npz"), **out)


def testnode_fixture():
    """tests/golden/reference_testnode.json: what the reference's own diagnostic node (HDK_TestGeometricMultigrid.cpp, compiled unmodified)
    prints for its MGPCG test on its own domains: iteration count and relative-residual history (ten digits)."""
    import json

    from oracle.bindings import TestNodeLib
    from tests.test_reference_testnode import CG_CASES, FIXTURE, node_cg

    node = TestNodeLib()
    out = {}
    for dom, n in CG_CASES:
        it, hist = node_cg(node, dom, n)
        out[f"{dom}{n}"] = {"iterations": it, "history": hist}
        print(dom, n, "iterations", it, "final", hist[-1])
    json.dump(out, open(FIXTURE, "w"), indent=1)


if __name__ == "__main__":
    main()
    node_fixture()
    testnode_fixture()
