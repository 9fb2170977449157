"""Python host mirror of the reference's solver interface, over the C ABI of include/gmg_b200.h.

The product is the CUDA library ``libgmg_b200.so`` next to this file; this module only marshals
numpy arrays into it through ctypes.  There is no CPU path: importing works anywhere (so the
build check can run on a CPU box), but creating a :class:`Context` without the library or without
a CUDA device raises.

Names follow the reference (rgoldade/GeometricMultigridPressureSolver):
``buildExpandedCellLabels`` / ``buildExpandedBoundaryWeights`` / ``setBoundaryCellLabels`` /
``buildCoarseCellLabels`` / ``buildBoundaryCells`` (HDK_GeometricMultigridOperators.h:121-157),
``GeometricMultigridPoissonSolver`` with ``applyVCycle`` / ``getMGLevels``
(HDK_GeometricMultigridPoissonSolver.h:10-53) and ``solveGeometricConjugateGradient``
(HDK_GeometricCGPoissonSolver.h:11-27).  Arrays are C-order numpy with shape (rz, ry, rx): x fastest,
like UT_VoxelArray.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgmg_b200.so")

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)

STATUS = {0: "GMG_OK", 1: "GMG_ERR_CUDA", 2: "GMG_ERR_INVALID", 3: "GMG_ERR_NO_ACTIVE", 4: "GMG_ERR_COARSE_SIZE", 5: "GMG_ERR_NOT_SPD", 6: "GMG_ERR_COMM"}

# every symbol include/gmg_b200.h declares (tests/test_abi.py checks the header against this and the .so)
ABI_SYMBOLS = [
    "gmg_last_error", "gmg_version", "gmg_ctx_create", "gmg_ctx_destroy", "gmg_ctx_synchronize", "gmg_ctx_shard", "gmg_nccl_unique_id", "gmg_ctx_rank", "gmg_shard_plan", "gmg_solver_shard_info", "gmg_comm_count", "gmg_comm_benchmark",
    "gmg_expand_dims", "gmg_expand_labels", "gmg_expand_weights", "gmg_set_boundary_labels", "gmg_coarsen_labels", "gmg_boundary_cells",
    "gmg_solver_default_options", "gmg_solver_create", "gmg_solver_create_u8", "gmg_solver_destroy", "gmg_solver_levels", "gmg_solver_level_res",
    "gmg_solver_get_labels", "gmg_solver_get_boundary_cells", "gmg_solver_active_cells", "gmg_solver_coarse_unknowns", "gmg_solver_setup_ms", "gmg_solver_transfer_cells", "gmg_transfer_plan", "gmg_gather_face_weights",
    "gmg_vcycle", "gmg_pcg", "gmg_pcg_from_zero", "gmg_grid_create", "gmg_grid_destroy", "gmg_grid_upload", "gmg_grid_download", "gmg_grid_zero", "gmg_grid_copy",
    "gmg_jacobi", "gmg_gauss_seidel", "gmg_boundary_jacobi", "gmg_apply", "gmg_residual", "gmg_restrict", "gmg_prolong_add", "gmg_dot", "gmg_norm2", "gmg_inf_norm",
    "gmg_axpy", "gmg_add_scaled", "gmg_scale", "gmg_vcycle_device", "gmg_pcg_device", "gmg_launch_count", "gmg_timer_begin", "gmg_timer_end",
    "gmg_build_material_labels", "gmg_build_valid_faces", "gmg_build_domain_labels", "gmg_build_boundary_weights", "gmg_build_rhs", "gmg_apply_old_pressure", "gmg_apply_solution_to_pressure", "gmg_apply_pressure_gradient",
    "gmg_profile_enable", "gmg_kernel_class_count", "gmg_kernel_class_name", "gmg_profile_get", "gmg_profile_reset", "gmg_profile_get_level",
]


def expand_dims(base_shape):
    """Ops.h:1340-1360: (expanded shape (z,y,x), offset (x,y,z), mgLevels) of a base grid of shape (z,y,x).  Host-only."""
    lib = load_library()
    eres, off, lv = (C.c_int64 * 3)(), (C.c_int64 * 3)(), C.c_int()
    _check(lib.gmg_expand_dims(_res(base_shape), eres, off, C.byref(lv)))
    return (int(eres[2]), int(eres[1]), int(eres[0])), np.array(list(off), dtype=np.int64), int(lv.value)


def transfer_plan(extents):
    """gmg_transfer_plan: extents (planes, 4) = per z-plane (x0, x1, y0, y1) of the active cells -> (groups (k, 6) = (z0, z1, x0, x1, y0, y1),
    cells they hold).  Host-only, needs no GPU."""
    lib = load_library()
    e = np.ascontiguousarray(extents, dtype=np.int32).reshape(-1, 4)
    groups = np.zeros((max(len(e), 1), 6), dtype=np.int32)
    n, cells = C.c_int(), C.c_int64()
    _check(lib.gmg_transfer_plan(e.ctypes.data_as(_i32p), len(e), groups.ctypes.data_as(_i32p), C.byref(n), C.byref(cells)))
    return groups[: n.value].copy(), int(cells.value)


def gather_face_weights(idx, pitch, plane, org, weights, bounds=None):
    """gmg_gather_face_weights: (6, count) face weights of the cells with storage indices idx in a box of row pitch `pitch`, plane
    size `plane` and expanded origin org (x, y, z); weights = the three expanded face-weight grids.  Host-only, needs no GPU."""
    lib = load_library()
    i = np.ascontiguousarray(idx, dtype=np.int32)
    w = [np.ascontiguousarray(a, dtype=np.float64) for a in weights]
    res = [w[0].shape[2] - 1, w[0].shape[1], w[0].shape[0]]
    out = np.zeros((6, len(i)), dtype=np.float64)
    b = (C.c_int64 * 6)(*[int(v) for v in bounds]) if bounds is not None else None
    _check(lib.gmg_gather_face_weights(i.ctypes.data_as(_i32p), C.c_int64(len(i)), int(pitch), C.c_int64(plane), (C.c_int32 * 3)(*[int(v) for v in org]),
                                       w[0].ctypes.data_as(_f64p), w[1].ctypes.data_as(_f64p), w[2].ctypes.data_as(_f64p),
                                       (C.c_int64 * 3)(*res), b, out.ctypes.data_as(_f64p)))
    return out


def shard_plan(level_planes, level_shift_z, level_cells, world, max_shard_levels=3, min_cells=1500000):
    """gmg_shard_plan: (number of sharded levels, cuts[level][rank] in storage planes).  Host-only, needs no GPU."""
    lib = load_library()
    n = len(level_planes)
    arr = lambda v: (C.c_int64 * n)(*[int(x) for x in v])
    cuts = (C.c_int64 * (n * (world + 1)))()
    S = C.c_int()
    _check(lib.gmg_shard_plan(arr(level_planes), arr(level_shift_z), arr(level_cells), n, int(world), int(max_shard_levels), C.c_int64(min_cells),
                              C.byref(S), cuts))
    return int(S.value), [[int(cuts[l * (world + 1) + k]) for k in range(world + 1)] for l in range(S.value)]


class GmgError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class SolverOptions(C.Structure):
    _fields_ = [
        ("use_gauss_seidel", C.c_int),
        ("print_stats", C.c_int),
        ("boundary_width", C.c_int),
        ("boundary_iterations", C.c_int),
        ("coarse_matrix_scale", C.c_double),
        ("box_lo", C.c_int64 * 3),
        ("box_hi", C.c_int64 * 3),
        ("operators_only", C.c_int),
        ("mixed_precision", C.c_int),
    ]


_lib = None


def load_library():
    """Load libgmg_b200.so (fails loudly: the product has no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmgError(1, f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a)")
        lib = C.CDLL(LIB_PATH)
        lib.gmg_last_error.restype = C.c_char_p
        lib.gmg_kernel_class_name.restype = C.c_char_p
        lib.gmg_solver_default_options.restype = None
        _lib = lib
    return _lib


def _check(st):
    if st != 0:
        raise GmgError(st, load_library().gmg_last_error().decode())


def _res(shape):
    return (C.c_int64 * 3)(int(shape[2]), int(shape[1]), int(shape[0]))


def _vec3(v):
    return (C.c_int64 * 3)(int(v[0]), int(v[1]), int(v[2]))


def face_shape(shape, axis):
    s = list(shape)
    s[2 - axis] += 1
    return tuple(s)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


class Context:
    """One CUDA device + stream (gmg_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.gmg_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.world = 0, 1

    def synchronize(self):
        _check(self.lib.gmg_ctx_synchronize(self.h))

    # ---- z-slab sharding: one process per GPU, the 128-byte NCCL id travels over the caller's own channel ----
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        _check(self.lib.gmg_nccl_unique_id(buf))
        return buf.raw

    def shard(self, rank: int, world: int, unique_id: bytes | None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(self.lib.gmg_ctx_shard(self.h, int(rank), int(world), buf))
        self.rank, self.world = int(rank), int(world)

    def shard_with_torch(self, dist):
        """Shard over an initialised torch.distributed group: rank 0 makes the NCCL id, everyone receives it."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.shard(rank, world, box[0])

    def comm_count(self) -> int:
        n = C.c_int64()
        _check(self.lib.gmg_comm_count(self.h, C.byref(n)))
        return int(n.value)

    def launch_count(self, reset=False) -> int:
        n = C.c_int64()
        _check(self.lib.gmg_launch_count(self.h, C.byref(n), int(reset)))
        return int(n.value)

    def timer_begin(self):
        _check(self.lib.gmg_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_double()
        _check(self.lib.gmg_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def profile_enable(self, on=True):
        _check(self.lib.gmg_profile_enable(self.h, int(on)))

    def profile_reset(self):
        _check(self.lib.gmg_profile_reset(self.h))

    def profile(self, fine_level_only=False):
        """{class name: (ms, launches, algorithmic bytes)} accumulated since the last reset."""
        out = {}
        for i in range(self.lib.gmg_kernel_class_count()):
            ms, n, by = C.c_double(), C.c_int64(), C.c_double()
            _check(self.lib.gmg_profile_get(self.h, i, int(fine_level_only), C.byref(ms), C.byref(n), C.byref(by)))
            out[self.lib.gmg_kernel_class_name(i).decode()] = (ms.value, int(n.value), by.value)
        return out

    def profile_by_level(self, levels):
        """{level: {class name: (ms, launches)}} for the classes that ran on that level."""
        out = {}
        for l in range(levels):
            row = {}
            for i in range(self.lib.gmg_kernel_class_count()):
                ms, n = C.c_double(), C.c_int64()
                _check(self.lib.gmg_profile_get_level(self.h, i, l, C.byref(ms), C.byref(n)))
                if n.value:
                    row[self.lib.gmg_kernel_class_name(i).decode()] = (ms.value, int(n.value))
            out[l] = row
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.gmg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- domain builders (HDK::GeometricMultigridOperators) ------------------------------------
    def buildExpandedCellLabels(self, base_labels):
        """Ops.h:1328-1456.  Returns (expanded labels, offset (x,y,z), mgLevels)."""
        b, bp = _i32(base_labels)
        eres, off, lv = (C.c_int64 * 3)(), (C.c_int64 * 3)(), C.c_int()
        _check(self.lib.gmg_expand_dims(_res(b.shape), eres, off, C.byref(lv)))
        out = np.empty((eres[2], eres[1], eres[0]), dtype=np.int32)
        _check(self.lib.gmg_expand_labels(self.h, bp, _res(b.shape), out.ctypes.data_as(_i32p), eres, off))
        return out, np.array(list(off), dtype=np.int64), int(lv.value)

    def buildExpandedBoundaryWeights(self, base_weights, base_shape, exp_shape, offset, axis):
        """Ops.h:1458-1572 for one axis."""
        w, wp = _f64(base_weights)
        out = np.empty(face_shape(exp_shape, axis), dtype=np.float64)
        _check(self.lib.gmg_expand_weights(self.h, wp, _res(base_shape), out.ctypes.data_as(_f64p), _res(exp_shape), _vec3(offset), int(axis)))
        return out

    def setBoundaryCellLabels(self, labels, weights, box=None, inplace=False):
        """Ops.h:1574-1644 (returns a new array; inplace=True rewrites `labels` itself, like the reference)."""
        l, lp = _i32(labels if inplace else np.array(labels, copy=True))
        ws = [_f64(w) for w in weights]
        lo = _vec3(box[0]) if box else None
        hi = _vec3(box[1]) if box else None
        _check(self.lib.gmg_set_boundary_labels(self.h, lp, _res(l.shape), ws[0][1], ws[1][1], ws[2][1], lo, hi))
        return l

    def buildCoarseCellLabels(self, fine_labels):
        """Ops.cpp:23-163."""
        f, fp = _i32(fine_labels)
        out = np.empty(tuple(s // 2 for s in f.shape), dtype=np.int32)
        _check(self.lib.gmg_coarsen_labels(self.h, fp, _res(f.shape), out.ctypes.data_as(_i32p)))
        return out

    def buildBoundaryCells(self, labels, width=3):
        """Ops.cpp:165-469.  Returns int64 (n,3) in the reference's order."""
        l, lp = _i32(labels)
        n = C.c_int64()
        _check(self.lib.gmg_boundary_cells(self.h, lp, _res(l.shape), int(width), None, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), dtype=np.int64)
        _check(self.lib.gmg_boundary_cells(self.h, lp, _res(l.shape), int(width), out.ctypes.data_as(_i64p), C.byref(n)))
        return out[: n.value]

    # ---- the steps either side of the solve (GFS.cpp:746-1131; base-grid fields, fpreal32 like SIM_RawField) --------------
    @staticmethod
    def _f32(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def buildMaterialCellLabels(self, liquidSurface, solidSurface, cutCell):
        """HDK_Utilities.cpp:87-148: SOLID (0) / LIQUID (1) / AIR (2) from the surface SDF, the solid SDF sampled at the cell centres
        and the three cut-cell weight fields."""
        ls, lsp = self._f32(liquidSurface)
        so, sop = self._f32(solidSurface)
        assert so.shape == ls.shape and all(np.shape(cutCell[a]) == face_shape(ls.shape, a) for a in range(3))
        kc, cp = self._field_ptrs(cutCell)
        out = np.empty(ls.shape, dtype=np.int32)
        _check(self.lib.gmg_build_material_labels(self.h, lsp, sop, cp, _res(ls.shape), out.ctypes.data_as(_i32p)))
        return out

    def buildValidFaces(self, material, cutCell, axis):
        """GFS.cpp:717-744 for one axis: 1 on faces with a cut-cell weight > 0 next to a LIQUID cell, else 0 (fpreal32)."""
        m, mp = _i32(material)
        cc, ccp = self._f32(cutCell)
        assert cc.shape == face_shape(m.shape, axis)
        out = np.empty(cc.shape, dtype=np.float32)
        _check(self.lib.gmg_build_valid_faces(self.h, mp, ccp, _res(m.shape), int(axis), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def buildMGDomainLabels(self, material):
        """GFS.cpp:746-793: LIQUID (1) -> INTERIOR, AIR (2) -> DIRICHLET, else EXTERIOR."""
        m, mp = _i32(material)
        out = np.empty(m.shape, dtype=np.int32)
        _check(self.lib.gmg_build_domain_labels(self.h, mp, _res(m.shape), out.ctypes.data_as(_i32p)))
        return out

    def buildMGBoundaryWeights(self, cutCell, liquidSurface, validFaces, domainLabels, axis):
        """GFS.cpp:796-865 for one axis."""
        l, lp = _i32(domainLabels)
        cc, ccp = self._f32(cutCell)
        ls, lsp = self._f32(liquidSurface)
        vf, vfp = self._f32(validFaces)
        assert cc.shape == face_shape(l.shape, axis) and vf.shape == cc.shape and ls.shape == l.shape
        out = np.empty(cc.shape, dtype=np.float64)
        _check(self.lib.gmg_build_boundary_weights(self.h, ccp, lsp, vfp, lp, _res(l.shape), int(axis), out.ctypes.data_as(_f64p)))
        return out

    def _field_ptrs(self, fields):
        keep = [self._f32(f) for f in fields]
        return keep, (C.POINTER(C.c_float) * 3)(*[k[1] for k in keep])

    def buildRHS(self, material, velocity, cutCell, exp_shape, offset, solidVelocity=None):
        """GFS.cpp:868-943: the expanded rhs grid (zero off the LIQUID cells)."""
        m, mp = _i32(material)
        kv, vp = self._field_ptrs(velocity)
        kc, cp = self._field_ptrs(cutCell)
        sp = None
        if solidVelocity is not None:
            ks, sp = self._field_ptrs(solidVelocity)
        rhs = np.zeros(exp_shape, dtype=np.float64)
        _check(self.lib.gmg_build_rhs(self.h, mp, vp, cp, sp, _res(m.shape), _res(exp_shape), _vec3(offset), rhs.ctypes.data_as(_f64p)))
        return rhs

    def applyOldPressure(self, pressure, material, exp_shape, offset):
        """GFS.cpp:946-997: the warm-start solution grid."""
        m, mp = _i32(material)
        p, pp = self._f32(pressure)
        x = np.zeros(exp_shape, dtype=np.float64)
        _check(self.lib.gmg_apply_old_pressure(self.h, pp, mp, _res(m.shape), _res(exp_shape), _vec3(offset), x.ctypes.data_as(_f64p)))
        return x

    def applySolutionToPressure(self, pressure, material, solution, offset):
        """GFS.cpp:1000-1047: returns the updated fpreal32 pressure field."""
        m, mp = _i32(material)
        p, pp = self._f32(np.array(pressure, dtype=np.float32, copy=True))
        x, xp = _f64(solution)
        _check(self.lib.gmg_apply_solution_to_pressure(self.h, pp, mp, xp, _res(m.shape), _res(x.shape), _vec3(offset)))
        return p

    def applyPressureGradient(self, velocity, liquidSurface, pressure, validFaces, material, axis):
        """GFS.cpp:1050-1131 for one axis: returns the updated fpreal32 face velocities."""
        m, mp = _i32(material)
        v, vp = self._f32(np.array(velocity, dtype=np.float32, copy=True))
        ls, lsp = self._f32(liquidSurface)
        p, pp = self._f32(pressure)
        vf, vfp = self._f32(validFaces)
        _check(self.lib.gmg_apply_pressure_gradient(self.h, vp, lsp, pp, vfp, mp, _res(m.shape), int(axis)))
        return v

    def buildExpandedDomainLazy(self, base_labels, base_weights):
        """buildExpandedDomain for large grids: the expanded arrays come from np.zeros and only the base box is ever written,
        so the 8x larger virtual grid costs no host memory (untouched pages are never materialised).  OUTSIDE the returned
        box the arrays hold zeros, not EXTERIOR -- they are only valid for calls that take the box hint (this library reads
        host arrays inside the hinted box only)."""
        bl = np.asarray(base_labels)
        shape, off, levels = expand_dims(bl.shape)
        nz, ny, nx = bl.shape
        box = (slice(int(off[2]), int(off[2]) + nz), slice(int(off[1]), int(off[1]) + ny), slice(int(off[0]), int(off[0]) + nx))
        labels = np.zeros(shape, dtype=np.int32)
        labels[box] = np.where(bl == 1, 1, np.where(bl == 0, 0, 2))  # Ops.h:1404-1453
        w = []
        for a in range(3):
            we = np.zeros(face_shape(shape, a), dtype=np.float64)
            fb = list(box)
            fb[2 - a] = slice(fb[2 - a].start, fb[2 - a].stop + 1)
            we[tuple(fb)] = np.maximum(np.asarray(base_weights[a], dtype=np.float64), 0.0)  # Ops.h:1488-1571
            w.append(we)
        hi = [int(off[a]) + bl.shape[2 - a] for a in range(3)]
        self.setBoundaryCellLabels(labels, w, box=(off, hi), inplace=True)
        return labels, w, off, levels, (off, hi)

    def buildExpandedDomain(self, base_labels, base_weights):
        """Test.cpp:170-204 buildExpandedDomain: labels, three weight grids, setBoundaryCellLabels."""
        labels, offset, levels = self.buildExpandedCellLabels(base_labels)
        w = [self.buildExpandedBoundaryWeights(base_weights[a], base_labels.shape, labels.shape, offset, a) for a in range(3)]
        hi = [int(offset[a]) + base_labels.shape[2 - a] for a in range(3)]
        labels = self.setBoundaryCellLabels(labels, w, box=(offset, hi))
        return labels, w, offset, levels


class Grid:
    """Device-resident fp64 vector grid of one solver level (gmg_grid)."""

    def __init__(self, solver: "GeometricMultigridPoissonSolver", level: int = 0, host=None):
        self.solver = solver
        self.level = level
        h = C.c_void_p()
        _check(solver.lib.gmg_grid_create(solver.h, int(level), C.byref(h)))
        self.h = h
        if host is not None:
            self.upload(host)

    @property
    def shape(self):
        return self.solver.level_shape(self.level)

    def upload(self, host):
        a, ap = _f64(host)
        assert a.shape == self.shape, (a.shape, self.shape)
        _check(self.solver.lib.gmg_grid_upload(self.h, ap))
        return self

    def download(self):
        out = np.empty(self.shape, dtype=np.float64)
        _check(self.solver.lib.gmg_grid_download(self.h, out.ctypes.data_as(_f64p)))
        return out

    def zero(self):
        _check(self.solver.lib.gmg_grid_zero(self.h))
        return self

    def copy_from(self, other: "Grid"):
        _check(self.solver.lib.gmg_grid_copy(self.h, other.h))
        return self

    def close(self):
        if getattr(self, "h", None) and getattr(self.solver, "h", None):
            self.solver.lib.gmg_grid_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GeometricMultigridPoissonSolver:
    """HDK::GeometricMultigridPoissonSolver (HDK_GeometricMultigridPoissonSolver.h:10-53) on the GPU."""

    def __init__(self, ctx: Context, initialCellLabels, boundaryWeights, mgLevels: int, useGaussSeidel: bool = False, doPrintStats: bool = False,
                 coarse_matrix_scale: float = 1.0, box=None, boundary_iterations: int = 3, boundary_width: int = 3, mixed_precision: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        # one-byte labels (uint8 / int8 arrays) go through gmg_solver_create_u8: a quarter of the host memory and PCIe traffic
        byte_labels = np.asarray(initialCellLabels).dtype.itemsize == 1
        if byte_labels:
            l = np.ascontiguousarray(initialCellLabels).view(np.uint8)
            lp = l.ctypes.data_as(C.c_void_p)
        else:
            l, lp = _i32(initialCellLabels)
        if boundaryWeights is None:  # the operators' `boundaryWeights == nullptr` form: weight 1 (Ops.h:237-248)
            ws = [(None, None)] * 3
        else:
            ws = [_f64(w) for w in boundaryWeights]
            for a in range(3):
                assert ws[a][0].shape == face_shape(l.shape, a), "boundary weight grid shape (MG.cpp:167-177)"
        opt = SolverOptions()
        self.lib.gmg_solver_default_options(C.byref(opt))
        opt.use_gauss_seidel = int(useGaussSeidel)
        opt.print_stats = int(doPrintStats)
        opt.coarse_matrix_scale = float(coarse_matrix_scale)
        opt.boundary_iterations = int(boundary_iterations)
        opt.boundary_width = int(boundary_width)
        opt.mixed_precision = int(mixed_precision)  # fp32 V-cycle inside the fp64 CG (README.md:34-35 TODO of the reference); not reference arithmetic
        if box is not None:
            for a in range(3):
                opt.box_lo[a] = int(box[0][a])
                opt.box_hi[a] = int(box[1][a])
        h = C.c_void_p()
        create = self.lib.gmg_solver_create_u8 if byte_labels else self.lib.gmg_solver_create
        _check(create(ctx.h, lp, _res(l.shape), ws[0][1], ws[1][1], ws[2][1], int(mgLevels), C.byref(opt), C.byref(h)))
        self.h = h
        self.shape = l.shape

    def getMGLevels(self) -> int:
        n = C.c_int()
        _check(self.lib.gmg_solver_levels(self.h, C.byref(n)))
        return int(n.value)

    def level_shape(self, level):
        r = (C.c_int64 * 3)()
        _check(self.lib.gmg_solver_level_res(self.h, int(level), r))
        return (int(r[2]), int(r[1]), int(r[0]))

    def level_labels(self, level):
        out = np.empty(self.level_shape(level), dtype=np.int32)
        _check(self.lib.gmg_solver_get_labels(self.h, int(level), out.ctypes.data_as(_i32p)))
        return out

    def level_boundary_cells(self, level):
        n = C.c_int64()
        _check(self.lib.gmg_solver_get_boundary_cells(self.h, int(level), None, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), dtype=np.int64)
        _check(self.lib.gmg_solver_get_boundary_cells(self.h, int(level), out.ctypes.data_as(_i64p), C.byref(n)))
        return out[: n.value]

    def active_cells(self, level=0) -> int:
        n = C.c_int64()
        _check(self.lib.gmg_solver_active_cells(self.h, int(level), C.byref(n)))
        return int(n.value)

    def shard_info(self, level=0):
        """(is this level a z-slab, expanded z range [lo, hi) of the owned planes, active cells stored on this rank)"""
        sh, lo, hi, act = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(self.lib.gmg_solver_shard_info(self.h, int(level), C.byref(sh), C.byref(lo), C.byref(hi), C.byref(act)))
        return bool(sh.value), int(lo.value), int(hi.value), int(act.value)

    def comm_benchmark(self, kind: int, level: int = 0, depth: int = 8, reps: int = 100) -> float:
        """ms per back-to-back communication operation (0 halo exchange, 1 replicated-level gather, 2 scalar all-reduce)."""
        ms = C.c_double()
        _check(self.lib.gmg_comm_benchmark(self.h, int(kind), int(level), int(depth), int(reps), C.byref(ms)))
        return float(ms.value)

    def coarse_unknowns(self) -> int:
        n = C.c_int64()
        _check(self.lib.gmg_solver_coarse_unknowns(self.h, C.byref(n)))
        return int(n.value)

    def transfer_cells(self):
        """(cells one host<->device transfer of a level-0 vector grid moves, number of 3D copies): only the bounding rectangles of the
        active cells of each z-plane travel."""
        n, k = C.c_int64(), C.c_int64()
        _check(self.lib.gmg_solver_transfer_cells(self.h, C.byref(n), C.byref(k)))
        return int(n.value), int(k.value)

    def setup_ms(self) -> float:
        ms = C.c_double()
        _check(self.lib.gmg_solver_setup_ms(self.h, C.byref(ms)))
        return float(ms.value)

    # ---- host-buffer entry points (what the reference's callers use) ----------------------------
    def applyVCycle(self, solutionVector, rhsVector, useInitialGuess: bool = False, inplace: bool = False):
        """MG.cpp:420-881; returns the new solution grid (inplace=True mutates solutionVector like the reference does)."""
        x, xp = _f64(solutionVector if inplace else np.array(solutionVector, copy=True))
        b, bp = _f64(rhsVector)
        _check(self.lib.gmg_vcycle(self.h, xp, bp, int(useInitialGuess)))
        return x

    @staticmethod
    def _precond(useMGPreconditioner):
        """The node's useMGPreconditioner toggle (GFS.cpp:428): True = multigrid V-cycle, "diagonal" (or 2) = its other branch,
        the diagonal preconditioner of GFS.cpp:485-618; False = none (plain CG, not a mode of the node)."""
        if isinstance(useMGPreconditioner, str):
            return {"mg": 1, "multigrid": 1, "diagonal": 2, "none": 0}[useMGPreconditioner]
        return int(useMGPreconditioner)

    def solveGeometricConjugateGradient(self, solutionGrid, rhsGrid, tolerance, maxIterations, useMGPreconditioner=True,
                                        inplace: bool = False, solutionIsZero: bool = False):
        """CG.h:11-207 with A = applyPoissonMatrix, M^-1 = applyVCycle (GFS.cpp:430-483) or the diagonal preconditioner
        (GFS.cpp:485-618, useMGPreconditioner="diagonal").
        Returns (solution, iterations printed by CG.h:198 or -1 on an early-out, relative-residual history).
        inplace=True writes the pressure into solutionGrid itself (as the reference does) instead of a copy.
        solutionIsZero=True: the caller declares solutionGrid constant zero on entry (the node's solutionGrid.constant(0) without a warm
        start, GFS.cpp:392-398; solutionGrid may then be None): it is not uploaded (gmg_pcg_from_zero)."""
        b, bp = _f64(rhsGrid)
        if solutionGrid is None:
            if not solutionIsZero:
                raise ValueError("solutionGrid is None without solutionIsZero")
            solutionGrid, inplace = np.zeros(b.shape, dtype=np.float64), True
        x, xp = _f64(solutionGrid if inplace else np.array(solutionGrid, copy=True))
        hist = np.zeros(int(maxIterations) + 2, dtype=np.float64)
        it, cnt = C.c_int(), C.c_int()
        fn = self.lib.gmg_pcg_from_zero if solutionIsZero else self.lib.gmg_pcg
        _check(fn(self.h, xp, bp, C.c_double(tolerance), int(maxIterations), self._precond(useMGPreconditioner), C.byref(it),
                  hist.ctypes.data_as(_f64p), len(hist), C.byref(cnt)))
        return x, int(it.value), hist[: cnt.value].copy()

    # ---- device-resident operators ----------------------------------------------------------------
    def grid(self, level=0, host=None) -> Grid:
        return Grid(self, level, host)

    def jacobiPoissonSmoother(self, x: Grid, b: Grid):
        _check(self.lib.gmg_jacobi(self.h, x.h, b.h))

    def tiledGaussSeidelPoissonSmoother(self, x: Grid, b: Grid, doSmoothOddTiles: bool, doSmoothForward: bool):
        _check(self.lib.gmg_gauss_seidel(self.h, x.h, b.h, int(doSmoothOddTiles), int(doSmoothForward)))

    def boundaryJacobiPoissonSmoother(self, x: Grid, b: Grid, sweeps=1):
        _check(self.lib.gmg_boundary_jacobi(self.h, x.h, b.h, int(sweeps)))

    def applyPoissonMatrix(self, dst: Grid, src: Grid):
        _check(self.lib.gmg_apply(self.h, dst.h, src.h))

    def computePoissonResidual(self, r: Grid, x: Grid, b: Grid):
        _check(self.lib.gmg_residual(self.h, r.h, x.h, b.h))

    def downsample(self, coarse: Grid, fine: Grid):
        _check(self.lib.gmg_restrict(self.h, coarse.h, fine.h))

    def upsampleAndAdd(self, fine: Grid, coarse: Grid):
        _check(self.lib.gmg_prolong_add(self.h, fine.h, coarse.h))

    def dotProduct(self, a: Grid, b: Grid) -> float:
        out = C.c_double()
        _check(self.lib.gmg_dot(self.h, a.h, b.h, C.byref(out)))
        return float(out.value)

    def squaredL2Norm(self, a: Grid) -> float:
        out = C.c_double()
        _check(self.lib.gmg_norm2(self.h, a.h, C.byref(out)))
        return float(out.value)

    def l2Norm(self, a: Grid) -> float:
        return float(np.sqrt(self.squaredL2Norm(a)))

    def infNorm(self, a: Grid) -> float:
        out = C.c_double()
        _check(self.lib.gmg_inf_norm(self.h, a.h, C.byref(out)))
        return float(out.value)

    def addToVector(self, dst: Grid, src: Grid, scale: float):
        _check(self.lib.gmg_axpy(self.h, dst.h, src.h, C.c_double(scale)))

    def addVectors(self, dst: Grid, a: Grid, v: Grid, scale: float):
        _check(self.lib.gmg_add_scaled(self.h, dst.h, a.h, v.h, C.c_double(scale)))

    def scaleVector(self, v: Grid, scale: float):
        _check(self.lib.gmg_scale(self.h, v.h, C.c_double(scale)))

    def applyVCycleDevice(self, x: Grid, b: Grid, useInitialGuess=False):
        _check(self.lib.gmg_vcycle_device(self.h, x.h, b.h, int(useInitialGuess)))

    def solveDevice(self, x: Grid, b: Grid, tolerance, maxIterations, useMGPreconditioner=True):
        hist = np.zeros(int(maxIterations) + 2, dtype=np.float64)
        it, cnt = C.c_int(), C.c_int()
        _check(self.lib.gmg_pcg_device(self.h, x.h, b.h, C.c_double(tolerance), int(maxIterations), self._precond(useMGPreconditioner), C.byref(it),
                                       hist.ctypes.data_as(_f64p), len(hist), C.byref(cnt)))
        return int(it.value), hist[: cnt.value].copy()

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.gmg_solver_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
