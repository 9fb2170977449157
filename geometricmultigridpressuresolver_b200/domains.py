"""Synthetic label / weight / right-hand-side generators for the BASELINE.json configs.

These produce the *inputs* of the hot path (base-grid cell labels, ghost-fluid face weights,
dx^2-scaled right-hand sides) exactly as SURVEY.md section 8(d) specifies them, so that the CUDA
product, the CPU oracle and the compiled reference all consume identical arrays.  Pure numpy,
no device code, no dependency on oracle/.

Conventions follow the reference: arrays are C-order with shape (rz, ry, rx) -- x fastest,
like UT_VoxelArray -- labels are int32 with the enum of
HDK_GeometricMultigridOperators.h:11, the face grid of axis a has one more entry along a, and
face (i,j,k) of axis 0 lies between cells (i-1,j,k) and (i,j,k).
"""
from __future__ import annotations

import numpy as np

INTERIOR, EXTERIOR, DIRICHLET, BOUNDARY = 0, 1, 2, 3


def face_shape(shape, axis):
    s = list(shape)
    s[2 - axis] += 1
    return tuple(s)


def _np_axis(axis: int) -> int:
    """numpy axis of the reference's axis (0=x,1=y,2=z)."""
    return 2 - axis


def ghost_fluid_theta(phi0: np.ndarray, phi1: np.ndarray) -> np.ndarray:
    """HDK_Utilities.h:25-42 computeGhostFluidWeight, vectorised."""
    theta = np.zeros_like(phi0, dtype=np.float64)
    both = (phi0 < 0) & (phi1 < 0)
    a = (phi0 < 0) & (phi1 >= 0)
    b = (phi0 >= 0) & (phi1 < 0)
    theta[both] = 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        theta[a] = (phi0 / (phi0 - phi1))[a]
        theta[b] = (phi1 / (phi1 - phi0))[b]
    return theta


def ghost_fluid_weights(labels: np.ndarray, phi: np.ndarray):
    """Face weights of SURVEY.md 8(d) config 1 (the rule of Test.cpp:406-461):
    0 if either cell is EXTERIOR (or outside the grid) or both are DIRICHLET; 1 between two
    INTERIOR cells; 1/clamp(theta,.01,1) on a liquid/air face."""
    weights = []
    for axis in range(3):
        na = _np_axis(axis)
        w = np.zeros(face_shape(labels.shape, axis), dtype=np.float64)
        n = labels.shape[na]
        back = [slice(None)] * 3
        fwd = [slice(None)] * 3
        face = [slice(None)] * 3
        back[na] = slice(0, n - 1)
        fwd[na] = slice(1, n)
        face[na] = slice(1, n)
        lb, lf = labels[tuple(back)], labels[tuple(fwd)]
        pb, pf = phi[tuple(back)].astype(np.float64), phi[tuple(fwd)].astype(np.float64)
        inner = np.zeros(lb.shape, dtype=np.float64)
        both_int = (lb == INTERIOR) & (lf == INTERIOR)
        mixed = ((lb == INTERIOR) & (lf == DIRICHLET)) | ((lb == DIRICHLET) & (lf == INTERIOR))
        inner[both_int] = 1.0
        theta = np.clip(ghost_fluid_theta(pb, pf), 0.01, 1.0)
        inner[mixed] = (1.0 / theta)[mixed]
        w[tuple(face)] = inner
        weights.append(w)
    return weights


def _cell_points(n, dtype=np.float64):
    """cell point p = dx*(i,j,k) (reference convention, Test.cpp:259)."""
    nz, ny, nx = n
    dx = 1.0 / max(nx, ny, nz)
    z, y, x = np.meshgrid(np.arange(nz, dtype=dtype) * dx, np.arange(ny, dtype=dtype) * dx, np.arange(nx, dtype=dtype) * dx, indexing="ij")
    return x, y, z, dx


def sphere_domain(n: int):
    """Configs 1/2: liquid sphere (r=0.3) in a solid box, air elsewhere.  Returns (labels, weights, dx)."""
    shape = (n, n, n)
    x, y, z, dx = _cell_points(shape)
    phi = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2) - 0.3
    labels = np.where(phi <= 0, INTERIOR, DIRICHLET).astype(np.int32)
    for a in range(3):
        sl = [slice(None)] * 3
        sl[a] = 0
        labels[tuple(sl)] = EXTERIOR
        sl[a] = n - 1
        labels[tuple(sl)] = EXTERIOR
    return labels, ghost_fluid_weights(labels, phi), dx


def simple_domain(n: int, band: int = 1):
    """Test.cpp:466-625 buildSimpleDomain: box with a `band`-cell DIRICHLET rim, INTERIOR inside,
    weight 1 on faces touching an INTERIOR cell and no EXTERIOR/outside cell."""
    labels = np.full((n, n, n), DIRICHLET, dtype=np.int32)
    labels[band : n - band, band : n - band, band : n - band] = INTERIOR
    weights = []
    for axis in range(3):
        na = _np_axis(axis)
        w = np.zeros(face_shape(labels.shape, axis), dtype=np.float64)
        back = [slice(None)] * 3
        fwd = [slice(None)] * 3
        face = [slice(None)] * 3
        back[na] = slice(0, n - 1)
        fwd[na] = slice(1, n)
        face[na] = slice(1, n)
        lb, lf = labels[tuple(back)], labels[tuple(fwd)]
        w[tuple(face)] = ((lb == INTERIOR) | (lf == INTERIOR)).astype(np.float64)
        weights.append(w)
    return labels, weights, 1.0 / n


def complex_domain(n: int):
    """Test.cpp:207-464 buildComplexDomain with useSolidSphere = false: air/liquid split by
    x - .5 + .25 sin(2 pi y + 4 pi z) (float32 points and SDF samples, as SIM_RawField stores them),
    domain-border faces closed, ghost-fluid weights on liquid/air faces."""
    shape = (n, n, n)
    dxf = np.float32(1.0 / n)
    idx = np.arange(n, dtype=np.float32) * dxf
    z, y, x = np.meshgrid(idx, idx, idx, indexing="ij")
    phi = (x.astype(np.float64) - 0.5 + 0.25 * np.sin(2.0 * np.pi * y.astype(np.float64) + 4.0 * np.pi * z.astype(np.float64))).astype(np.float32)
    weights = [np.ones(face_shape(shape, a), dtype=np.float64) for a in range(3)]
    for axis in range(3):
        na = _np_axis(axis)
        sl = [slice(None)] * 3
        sl[na] = 0
        weights[axis][tuple(sl)] = 0
        sl[na] = n
        weights[axis][tuple(sl)] = 0
    # every cell of an n>=2 grid keeps at least one open face -> no EXTERIOR cells
    labels = np.where(phi > 0, DIRICHLET, INTERIOR).astype(np.int32)
    for axis in range(3):
        na = _np_axis(axis)
        back = [slice(None)] * 3
        fwd = [slice(None)] * 3
        face = [slice(None)] * 3
        back[na] = slice(0, n - 1)
        fwd[na] = slice(1, n)
        face[na] = slice(1, n)
        lb, lf = labels[tuple(back)], labels[tuple(fwd)]
        pb, pf = phi[tuple(back)].astype(np.float64), phi[tuple(fwd)].astype(np.float64)
        inner = weights[axis][tuple(face)]
        both_dir = (lb == DIRICHLET) & (lf == DIRICHLET)
        mixed = (lb == DIRICHLET) ^ (lf == DIRICHLET)
        theta = np.clip(ghost_fluid_theta(pb, pf), 0.01, 1.0)
        inner[both_dir] = 0
        inner[mixed] = (inner / theta)[mixed]
        weights[axis][tuple(face)] = inner
    return labels, weights, float(dxf)


def flipsplash_domain(n: int, shape=None):
    """Config 3: flipSplash-shaped tank -- five solid walls (x, z sides and floor y=0), open top,
    pool below y = 0.25 N, a 3x3 grid of liquid blobs of radius 0.06 N at y = 0.65 N.
    `shape` = (nz, ny, nx) overrides the cube (e.g. (300, 200, 300) for the scene-true variant)."""
    shape = (n, n, n) if shape is None else tuple(shape)
    nz, ny, nx = shape
    N = float(max(shape))
    k, j, i = np.meshgrid(np.arange(nz, dtype=np.float64), np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing="ij")
    phi = j - 0.25 * ny  # pool plane: negative below the water level
    r = 0.06 * N
    for cx in (nx / 3.0, nx / 2.0, 2.0 * nx / 3.0):
        for cz in (nz / 3.0, nz / 2.0, 2.0 * nz / 3.0):
            phi = np.minimum(phi, np.sqrt((i - cx) ** 2 + (j - 0.65 * ny) ** 2 + (k - cz) ** 2) - r)
    phi /= N
    labels = np.where(phi <= 0, INTERIOR, DIRICHLET).astype(np.int32)
    labels[:, :, 0] = EXTERIOR
    labels[:, :, nx - 1] = EXTERIOR
    labels[0, :, :] = EXTERIOR
    labels[nz - 1, :, :] = EXTERIOR
    labels[:, 0, :] = EXTERIOR
    return labels, ghost_fluid_weights(labels, phi), 1.0 / N


def liquid_box_domain(n: int):
    """Config 4: solid shell, liquid inside, one DIRICHLET layer y = N-2 so the system is non-singular
    (SURVEY.md fact 9).  Weights are 1 between active cells and on liquid/air faces, 0 otherwise."""
    labels = np.full((n, n, n), EXTERIOR, dtype=np.int32)
    labels[1 : n - 1, 1 : n - 1, 1 : n - 1] = INTERIOR
    labels[1 : n - 1, n - 2, 1 : n - 1] = DIRICHLET
    weights = []
    for axis in range(3):
        na = _np_axis(axis)
        w = np.zeros(face_shape(labels.shape, axis), dtype=np.float64)
        back = [slice(None)] * 3
        fwd = [slice(None)] * 3
        face = [slice(None)] * 3
        back[na] = slice(0, n - 1)
        fwd[na] = slice(1, n)
        face[na] = slice(1, n)
        lb, lf = labels[tuple(back)], labels[tuple(fwd)]
        ok = (lb != EXTERIOR) & (lf != EXTERIOR) & ((lb == INTERIOR) | (lf == INTERIOR))
        w[tuple(face)] = ok.astype(np.float64)
        weights.append(w)
    return labels, weights, 1.0 / n


def _unit_weights(labels):
    """Face weights 1 between two non-EXTERIOR cells of which at least one is liquid, 0 otherwise (the rule of liquid_box_domain)."""
    n = labels.shape
    weights = []
    for axis in range(3):
        na = _np_axis(axis)
        w = np.zeros(face_shape(labels.shape, axis), dtype=np.float64)
        back = [slice(None)] * 3
        fwd = [slice(None)] * 3
        face = [slice(None)] * 3
        back[na] = slice(0, n[na] - 1)
        fwd[na] = slice(1, n[na])
        face[na] = slice(1, n[na])
        lb, lf = labels[tuple(back)], labels[tuple(fwd)]
        ok = (lb != EXTERIOR) & (lf != EXTERIOR) & ((lb == INTERIOR) | (lf == INTERIOR))
        w[tuple(face)] = ok.astype(np.float64)
        weights.append(w)
    return weights


def half_liquid_box_domain(n: int):
    """A solid box whose upper z half is liquid and lower z half air: z-slab sharding leaves the lower ranks WITHOUT any active
    cell (their reductions must still deliver 0 and keep the CG scalars in step -- the N = 8 failure of round 1)."""
    labels = np.full((n, n, n), EXTERIOR, dtype=np.int32)
    labels[1 : n - 1, 1 : n - 1, 1 : n - 1] = INTERIOR
    labels[1 : n // 2 + 3, 1 : n - 1, 1 : n - 1] = DIRICHLET
    return labels, _unit_weights(labels), 1.0 / n


def narrow_band_domain(n: int, thickness: int = 12):
    """Config 5: thin liquid sheet over solid terrain h(x,z) = N(0.5 + 0.1 sin(2 pi x/N) sin(2 pi z/N));
    y < h-thickness solid, h-thickness <= y <= h liquid, above air; x/z walls solid."""
    shape = (n, n, n)
    k, j, i = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij", sparse=True)
    h = n * (0.5 + 0.1 * np.sin(2 * np.pi * i / n) * np.sin(2 * np.pi * k / n))
    phi = np.broadcast_to((j - h) / n, shape)
    labels = np.where(j > h, DIRICHLET, np.where(j < h - thickness, EXTERIOR, INTERIOR)).astype(np.int32)
    labels = np.ascontiguousarray(np.broadcast_to(labels, shape)).copy()
    labels[:, :, 0] = EXTERIOR
    labels[:, :, n - 1] = EXTERIOR
    labels[0, :, :] = EXTERIOR
    labels[n - 1, :, :] = EXTERIOR
    labels[:, 0, :] = EXTERIOR
    return labels, ghost_fluid_weights(labels, np.ascontiguousarray(phi)), 1.0 / n


# ---------------------------------------------------------------------------- right-hand sides

def active_mask(labels: np.ndarray) -> np.ndarray:
    return (labels == INTERIOR) | (labels == BOUNDARY)


def delta_rhs(exp_labels: np.ndarray, centre_xyz, dx: float, amplitude: float = 1000.0) -> np.ndarray:
    """3x3x3 block of `amplitude` centred at expanded cell `centre_xyz`, then *dx^2 on active cells
    (Test.cpp:727-742 delta, :793-794 scaleVector)."""
    b = np.zeros(exp_labels.shape, dtype=np.float64)
    cx, cy, cz = [int(v) for v in centre_xyz]
    b[cz - 1 : cz + 2, cy - 1 : cy + 2, cx - 1 : cx + 2] = amplitude
    m = active_mask(exp_labels)
    b[m] *= dx * dx
    b[~m] = 0.0  # keep the vector-grid invariant (the reference's operators ignore these cells anyway)
    return b


def random_rhs(exp_labels: np.ndarray, dx: float, seed: int = 12345) -> np.ndarray:
    """seeded uniform(0,1)*dx^2 on active cells, exactly 0 elsewhere."""
    rng = np.random.default_rng(seed)
    b = np.zeros(exp_labels.shape, dtype=np.float64)
    m = active_mask(exp_labels)
    b[m] = rng.random(int(m.sum())) * dx * dx
    return b


def random_active(exp_labels: np.ndarray, seed: int, scale: float = 1.0) -> np.ndarray:
    """random values in [-scale, scale) on active cells, 0 elsewhere (the vector-grid invariant)."""
    rng = np.random.default_rng(seed)
    v = np.zeros(exp_labels.shape, dtype=np.float64)
    m = active_mask(exp_labels)
    v[m] = (rng.random(int(m.sum())) * 2.0 - 1.0) * scale
    return v


DOMAINS = {
    "sphere": sphere_domain,
    "simple": simple_domain,
    "complex": complex_domain,
    "flipsplash": flipsplash_domain,
    "liquid_box": liquid_box_domain,
    "narrow_band": narrow_band_domain,
    "half_box": half_liquid_box_domain,
}
