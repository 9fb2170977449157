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


if __name__ == "__main__":
    main()
