"""Shared helpers of the test-suite: golden-case loading and input regeneration."""
import hashlib
import os

import numpy as np

from geometricmultigridpressuresolver_b200 import domains as D

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror tests/golden/make_golden.py:CASES
CASES = {
    "simple16": ("simple", 16, {}),
    "complex16": ("complex", 16, {}),
    "sphere24": ("sphere", 24, {}),
    "flipsplash_24x16x24": ("flipsplash", 24, {"shape": (24, 16, 24)}),
    "liquid_box16": ("liquid_box", 16, {}),
    "narrow_band32": ("narrow_band", 32, {"thickness": 4}),
    "sphere32": ("sphere", 32, {}),
    "sphere64": ("sphere", 64, {}),
}
FULL_CASES = [c for c in CASES if c != "sphere64"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def crop(a, off, base_shape):
    return a[off[2] : off[2] + base_shape[0], off[1] : off[1] + base_shape[1], off[0] : off[0] + base_shape[2]]


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def base_inputs(name):
    dom, n, kw = CASES[name]
    return D.DOMAINS[dom](n, **kw)


def rhs_for(labels, off, base_shape, dx):
    c = [int(off[0]) + base_shape[2] // 2, int(off[1]) + base_shape[1] // 2, int(off[2]) + base_shape[0] // 2]
    b = D.delta_rhs(labels, c, dx)
    if not b.any():
        b = D.random_rhs(labels, dx)
    return b


def relerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


# ---- fields of the front-end fixture tests/golden/frontend_fields.npz (tests/golden/make_golden_frontend.py) ----------------
FRONTEND_CASES = {"tank24": ((24, 24, 24), 11), "slab_40x18x33": ((40, 18, 33), 12), "thin_17x1x35": ((17, 1, 35), 13)}


def frontend_fields(name):
    """Seeded surface SDF, solid SDF (both signs) and cut-cell weights (closed, open, fractional; one solid block with every face
    closed) of a front-end fixture case: (phi, solid, [cut_x, cut_y, cut_z]), all float32."""
    shape, seed = FRONTEND_CASES[name]
    rng = np.random.default_rng(seed)
    phi = rng.random(shape).astype(np.float32) - np.float32(0.45)
    solid = rng.random(shape).astype(np.float32) - np.float32(0.5)
    cut = []
    for a in range(3):
        fs = D.face_shape(shape, a)
        cut.append((rng.random(fs) < 0.65).astype(np.float32) * (rng.random(fs).astype(np.float32) * np.float32(0.95) + np.float32(0.05)))
    blk = tuple(slice(0, max(1, s // 2)) for s in shape)
    for a in range(3):
        sl = list(blk)
        sl[2 - a] = slice(0, blk[2 - a].stop + 1)
        cut[a][tuple(sl)] = 0.0
    return phi, solid, cut
