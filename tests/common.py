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
