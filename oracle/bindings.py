"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``RefLib``  -> oracle/_ref/libgmg_ref.so : the reference's own sources compiled unmodified
  against oracle/shim (built by ``make -C oracle ref`` where /root/reference exists).
* ``PortLib`` -> oracle/libgmg_oracle.so   : the plain-C restatement (oracle/gmg_oracle.c).

Both expose the same Python surface so tests can swap them.  Arrays are numpy, C-order with
shape (rz, ry, rx), i.e. x-fastest like UT_VoxelArray; ``res`` passed to C is (rx, ry, rz).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libgmg_ref.so")
REF_DBG_SO = os.path.join(HERE, "_ref", "libgmg_ref_dbg.so")
REF_TESTNODE_SO = os.path.join(HERE, "_ref", "libgmg_ref_testnode.so")
PORT_SO = os.path.join(HERE, "libgmg_oracle.so")

_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(ref: bool = True) -> None:
    """Compile the checkers (port always; the reference bridge only where /root/reference exists)."""
    subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference/Source"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def _res(a: np.ndarray):
    return (C.c_int64 * 3)(a.shape[2], a.shape[1], a.shape[0])


def _res_t(t):
    return (C.c_int64 * 3)(*[int(v) for v in t])


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


def _wptrs(w):
    if w is None:
        return [None, None, None], [None, None, None]
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in w]
    return keep, [a.ctypes.data_as(_f64p) for a in keep]


def face_shape(shape, axis):
    """numpy shape (rz,ry,rx) of the face grid along `axis` (0=x)."""
    s = list(shape)
    s[2 - axis] += 1
    return tuple(s)


class _Base:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path)

    def fn(self, name, restype=C.c_int):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- shared operator wrappers (identical C shapes in both libraries) ----
    def set_boundary_labels(self, labels, w):
        l, lp = _i32(np.array(labels, copy=True))
        keep, wp = _wptrs(w)
        self.fn("set_boundary_labels", None)(lp, _res(l), *wp)
        return l

    def coarsen_labels(self, fine):
        f, fp = _i32(fine)
        out = np.empty(tuple(s // 2 for s in f.shape), dtype=np.int32)
        self.fn("coarsen_labels", None)(fp, _res(f), out.ctypes.data_as(_i32p))
        return out

    def unit_test_boundary_cells(self, labels, w=None):
        l, lp = _i32(labels)
        keep, wp = _wptrs(w)
        return bool(self.fn("unit_test_boundary_cells")(lp, _res(l), *wp))

    def unit_test_exterior_cells(self, labels):
        l, lp = _i32(labels)
        return bool(self.fn("unit_test_exterior_cells")(lp, _res(l)))

    def unit_test_coarsening(self, coarse, fine):
        c, cp = _i32(coarse)
        f, fp = _i32(fine)
        return bool(self.fn("unit_test_coarsening")(cp, fp, _res(f)))

    def jacobi(self, x, b, labels, w=None):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        l, lp = _i32(labels)
        keep, wp = _wptrs(w)
        self.fn("jacobi", None)(xp, bp, lp, _res(l), *wp)
        return x

    def gauss_seidel(self, x, b, labels, odd, forward, w=None):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        l, lp = _i32(labels)
        keep, wp = _wptrs(w)
        self.fn("gauss_seidel", None)(xp, bp, lp, _res(l), int(odd), int(forward), *wp)
        return x

    def boundary_jacobi(self, x, b, labels, cells, sweeps=1, w=None):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        l, lp = _i32(labels)
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        keep, wp = _wptrs(w)
        self.fn("boundary_jacobi", None)(xp, bp, lp, _res(l), cells.ctypes.data_as(_i64p), C.c_int64(len(cells)), int(sweeps), *wp)
        return x

    def apply(self, src, labels, w=None, dst=None):
        s, sp = _f64(src)
        d, dp = _f64(np.zeros_like(s) if dst is None else np.array(dst, copy=True))
        l, lp = _i32(labels)
        keep, wp = _wptrs(w)
        self.fn("apply", None)(dp, sp, lp, _res(l), *wp)
        return d

    def residual(self, x, b, labels, w=None):
        x, xp = _f64(x)
        b, bp = _f64(b)
        l, lp = _i32(labels)
        r = np.empty_like(x)
        keep, wp = _wptrs(w)
        self.fn("residual", None)(r.ctypes.data_as(_f64p), xp, bp, lp, _res(l), *wp)
        return r

    def upsample_add(self, fine, coarse, fine_labels, coarse_labels):
        f, fp = _f64(np.array(fine, copy=True))
        c, cp = _f64(coarse)
        fl, flp = _i32(fine_labels)
        if self.prefix == "ref_":
            cl, clp = _i32(coarse_labels)
            self.fn("upsample_add", None)(fp, cp, flp, clp, _res(f))
        else:
            self.fn("upsample_add", None)(fp, cp, flp, _res(f))
        return f

    def downsample(self, fine, coarse_labels, fine_labels):
        f, fp = _f64(fine)
        cl, clp = _i32(coarse_labels)
        out = np.empty(cl.shape, dtype=np.float64)
        if self.prefix == "ref_":
            fl, flp = _i32(fine_labels)
            self.fn("downsample", None)(out.ctypes.data_as(_f64p), fp, clp, flp, _res(f))
        else:
            self.fn("downsample", None)(out.ctypes.data_as(_f64p), fp, clp, _res(f))
        return out

    def dot(self, a, b, labels):
        a, ap = _f64(a)
        b, bp = _f64(b)
        l, lp = _i32(labels)
        return float(self.fn("dot", C.c_double)(ap, bp, lp, _res(l)))

    def norm2(self, a, labels):
        a, ap = _f64(a)
        l, lp = _i32(labels)
        return float(self.fn("norm2", C.c_double)(ap, lp, _res(l)))

    def inf_norm(self, a, labels):
        a, ap = _f64(a)
        l, lp = _i32(labels)
        return float(self.fn("inf_norm", C.c_double)(ap, lp, _res(l)))

    def axpy(self, dst, src, s, labels):
        d, dp = _f64(np.array(dst, copy=True))
        sr, sp = _f64(src)
        l, lp = _i32(labels)
        self.fn("axpy", None)(dp, sp, C.c_double(s), lp, _res(l))
        return d

    def add_scaled(self, a, v, s, labels, dst=None):
        a, ap = _f64(a)
        v, vp = _f64(v)
        d, dp = _f64(np.zeros_like(a) if dst is None else np.array(dst, copy=True))
        l, lp = _i32(labels)
        self.fn("add_scaled", None)(dp, ap, vp, C.c_double(s), lp, _res(l))
        return d

    def scale(self, v, s, labels):
        v, vp = _f64(np.array(v, copy=True))
        l, lp = _i32(labels)
        self.fn("scale", None)(vp, C.c_double(s), lp, _res(l))
        return v

    def expand_domain(self, base_labels, base_weights):
        """buildExpandedDomain of Test.cpp:170-204: labels + 3 weights + setBoundaryCellLabels."""
        labels, offset, levels = self.expand_labels(base_labels)
        w = self.expand_weights(base_weights, base_labels.shape, labels, offset)
        labels = self.set_boundary_labels(labels, w)
        return labels, w, offset, levels


class RefLib(_Base):
    prefix = "ref_"

    def __init__(self, debug=False):
        super().__init__(REF_DBG_SO if debug else REF_SO)

    def threads(self):
        return int(self.fn("threads")())

    def set_threads(self, n=None):
        """Use n (default: every host core) OpenMP threads, whatever OMP_NUM_THREADS the launcher exported."""
        self.fn("set_threads", None)(int(n or os.cpu_count() or 1))
        return self.threads()

    def build_material_labels(self, liquid_surface, solid_at_centres, cut_cell):
        """HDK::Utilities::buildMaterialCellLabels -- the reference's own HDK_Utilities.cpp, compiled unmodified."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        ls, so, cc = f32(liquid_surface), f32(solid_at_centres), [f32(c) for c in cut_cell]
        fp = C.POINTER(C.c_float)
        out = np.empty(ls.shape, dtype=np.int32)
        self.fn("build_material_labels", None)(ls.ctypes.data_as(fp), so.ctypes.data_as(fp), cc[0].ctypes.data_as(fp), cc[1].ctypes.data_as(fp), cc[2].ctypes.data_as(fp),
                                               _res(ls), out.ctypes.data_as(_i32p))
        return out

    # ---- the node's own source, HDK_GeometricFreeSurfacePressureSolver.cpp compiled unmodified (shim/hdk_node_shim.h) ----
    @staticmethod
    def _f32c(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def _f32x3(self, fields, copy=False):
        keep = [self._f32c(np.array(f, dtype=np.float32, copy=True) if copy else f) for f in fields]
        return [k[0] for k in keep], (C.POINTER(C.c_float) * 3)(*[k[1] for k in keep])

    def node_domain_labels(self, material):
        m, mp = _i32(material)
        out = np.empty(m.shape, dtype=np.int32)
        self.fn("node_domain_labels", None)(mp, _res(m), out.ctypes.data_as(_i32p))
        return out

    def node_boundary_weights(self, cut_cell, liquid_surface, valid_faces, material, domain_labels, axis):
        m, mp = _i32(material)
        l, lp = _i32(domain_labels)
        cc, ccp = self._f32c(cut_cell)
        ls, lsp = self._f32c(liquid_surface)
        vf, vfp = self._f32c(valid_faces)
        out = np.empty(cc.shape, dtype=np.float64)
        self.fn("node_boundary_weights", None)(ccp, lsp, vfp, mp, lp, _res(m), int(axis), out.ctypes.data_as(_f64p))
        return out

    def node_rhs(self, material, velocity, cut_cell, exp_labels, offset, solid_velocity=None):
        m, mp = _i32(material)
        el, elp = _i32(exp_labels)
        kv, vp = self._f32x3(velocity)
        kc, cp = self._f32x3(cut_cell)
        sp = None
        if solid_velocity is not None:
            ks, sp = self._f32x3(solid_velocity)
        rhs = np.zeros(el.shape, dtype=np.float64)
        self.fn("node_rhs", None)(mp, vp, cp, sp, _res(m), elp, _res(el), _res_t(offset), rhs.ctypes.data_as(_f64p))
        return rhs

    def node_old_pressure(self, pressure, material, exp_labels, offset):
        m, mp = _i32(material)
        el, elp = _i32(exp_labels)
        p, pp = self._f32c(pressure)
        x = np.zeros(el.shape, dtype=np.float64)
        self.fn("node_old_pressure", None)(pp, mp, _res(m), elp, _res(el), _res_t(offset), x.ctypes.data_as(_f64p))
        return x

    def node_solution_to_pressure(self, pressure, material, solution, exp_labels, offset):
        m, mp = _i32(material)
        el, elp = _i32(exp_labels)
        p, pp = self._f32c(np.array(pressure, dtype=np.float32, copy=True))
        x, xp = _f64(solution)
        self.fn("node_solution_to_pressure", None)(pp, mp, xp, _res(m), elp, _res(el), _res_t(offset))
        return p

    def node_pressure_gradient(self, velocity, cut_cell, liquid_surface, pressure, valid_faces, material, axis):
        m, mp = _i32(material)
        v, vp = self._f32c(np.array(velocity, dtype=np.float32, copy=True))
        cc, ccp = self._f32c(cut_cell)
        ls, lsp = self._f32c(liquid_surface)
        p, pp = self._f32c(pressure)
        vf, vfp = self._f32c(valid_faces)
        self.fn("node_pressure_gradient", None)(vp, ccp, lsp, pp, vfp, mp, _res(m), int(axis))
        return v

    def node_solve(self, liquid_surface, velocity, cut_cell, solid_surface=None, solid_velocity=None, pressure=None, density=1000.0, tolerance=1e-5,
                   max_iterations=2500, use_mg_preconditioner=True, use_old_pressure=False):
        """HDK_GeometricFreeSurfacePressureSolver::solveGasSubclass, the reference's whole pressure projection on its production wiring.
        Returns (ok, pressure, [velocity x3], [validFaces x3], log)."""
        ls, lsp = self._f32c(liquid_surface)
        vel, vp = self._f32x3(velocity, copy=True)
        kc, cp = self._f32x3(cut_cell)
        sop = svp = None
        if solid_surface is not None:
            so, sop = self._f32c(solid_surface)
        if solid_velocity is not None:
            ksv, svp = self._f32x3(solid_velocity)
        pr, prp = self._f32c(np.zeros(ls.shape, np.float32) if pressure is None else np.array(pressure, dtype=np.float32, copy=True))
        valid, vfp = self._f32x3([np.zeros(v.shape, np.float32) for v in vel])
        log = C.create_string_buffer(1 << 16)
        ok = self.fn("node_solve", C.c_int)(lsp, sop, vp, cp, svp, prp, vfp, C.c_float(density), _res(ls), C.c_double(tolerance), int(max_iterations),
                                            int(bool(use_mg_preconditioner)), int(bool(use_old_pressure)), log, len(log))
        return bool(ok), pr, vel, valid, log.value.decode(errors="replace")

    def build_valid_faces(self, material, cut_cell, axis):
        """findOccupiedFaceTiles + uncompressTiles + classifyValidFaces (the reference's own templates) in the order of GFS.cpp:717-744."""
        m, mp = _i32(material)
        cc = np.ascontiguousarray(cut_cell, dtype=np.float32)
        out = np.empty(cc.shape, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        self.fn("build_valid_faces", None)(mp, cc.ctypes.data_as(fp), _res(m), int(axis), out.ctypes.data_as(fp))
        return out

    def expand_labels(self, base):
        b, bp = _i32(base)
        out = _i32p()
        ores = (C.c_int64 * 3)()
        off = (C.c_int64 * 3)()
        lv = C.c_int()
        self.fn("expand_labels")(bp, _res(b), C.byref(out), ores, off, C.byref(lv))
        shape = (ores[2], ores[1], ores[0])
        arr = np.ctypeslib.as_array(out, shape=shape).copy()
        self.fn("free", None)(out)
        return arr, np.array(list(off), dtype=np.int64), int(lv.value)

    def expand_weights(self, base_w, base_shape, exp_labels, offset):
        l, lp = _i32(exp_labels)
        res_b = _res_t(base_shape[::-1])
        outs = []
        for axis in range(3):
            bw, bwp = _f64(base_w[axis])
            out = np.empty(face_shape(l.shape, axis), dtype=np.float64)
            self.fn("expand_weights")(bwp, res_b, lp, _res(l), _res_t(offset), axis, out.ctypes.data_as(_f64p))
            outs.append(out)
        return outs

    def boundary_cells(self, labels, width=3):
        l, lp = _i32(labels)
        out = _i64p()
        n = C.c_int64()
        self.fn("boundary_cells")(lp, _res(l), int(width), C.byref(out), C.byref(n))
        arr = np.ctypeslib.as_array(out, shape=(max(n.value, 1), 3))[: n.value].copy()
        self.fn("free", None)(out)
        return arr

    def solver(self, labels, w, levels, use_gs=False, coarse_scale=1.0):
        assert coarse_scale == 1.0 or int(os.environ.get("GMG_SHIM_JOBS", "1")) == int(coarse_scale)
        return _RefSolver(self, labels, w, levels, use_gs)


class _RefSolver:
    def __init__(self, lib, labels, w, levels, use_gs):
        self.lib = lib
        self.labels, lp = _i32(labels)
        keep, wp = _wptrs(w)
        f = lib.fn("solver_create", C.c_void_p)
        self.h = C.c_void_p(f(lp, _res(self.labels), *wp, int(levels), int(use_gs)))

    @property
    def levels(self):
        return int(self.lib.fn("solver_levels")(self.h))

    @property
    def setup_seconds(self):
        return float(self.lib.fn("solver_setup_seconds", C.c_double)(self.h))

    def vcycle(self, x, b, use_initial_guess=False, repeats=1):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        self.last_seconds = float(self.lib.fn("solver_vcycle", C.c_double)(self.h, xp, bp, int(use_initial_guess), int(repeats)))
        return x

    def pcg(self, x, b, tol, max_it, diagonal=False):
        """CG.h:11-207 with the multigrid V-cycle (GFS.cpp:468-483) or, diagonal=True, the diagonal preconditioner of
        GFS.cpp:485-618."""
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        hist = np.zeros(max_it + 2, dtype=np.float64)
        cnt = C.c_int()
        secs = C.c_double()
        it = self.lib.fn("pcg_diag" if diagonal else "pcg")(self.h, xp, bp, C.c_double(tol), int(max_it), hist.ctypes.data_as(_f64p), len(hist),
                                                            C.byref(cnt), C.byref(secs))
        self.last_seconds = secs.value
        return x, int(it), hist[: cnt.value].copy()

    def close(self):
        if self.h:
            self.lib.fn("solver_destroy", None)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PortLib(_Base):
    prefix = "orc_"

    def __init__(self):
        super().__init__(PORT_SO)

    # ---- restatement of the steps either side of the solve (GFS.cpp:717-1131; pinned to the node's own source by tests/test_node_reference.py) -------
    @staticmethod
    def _f32(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def _ptrs3(self, fields):
        keep = [self._f32(f) for f in fields]
        return keep, (C.POINTER(C.c_float) * 3)(*[k[1] for k in keep])

    def build_material_labels(self, liquid_surface, solid_at_centres, cut_cell):
        ls, lsp = self._f32(liquid_surface)
        so, sop = self._f32(solid_at_centres)
        kc, cp = self._ptrs3(cut_cell)
        out = np.empty(ls.shape, dtype=np.int32)
        self.fn("build_material_labels", None)(lsp, sop, cp, _res(ls), out.ctypes.data_as(_i32p))
        return out

    def build_valid_faces(self, material, cut_cell, axis):
        m, mp = _i32(material)
        cc, ccp = self._f32(cut_cell)
        out = np.empty(cc.shape, dtype=np.float32)
        self.fn("build_valid_faces", None)(mp, ccp, _res(m), int(axis), out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def build_domain_labels(self, material):
        m, mp = _i32(material)
        out = np.empty(m.shape, dtype=np.int32)
        self.fn("build_domain_labels", None)(mp, _res(m), out.ctypes.data_as(_i32p))
        return out

    def build_boundary_weights(self, cut_cell, liquid_surface, valid_faces, domain_labels, axis):
        l, lp = _i32(domain_labels)
        cc, ccp = self._f32(cut_cell)
        ls, lsp = self._f32(liquid_surface)
        vf, vfp = self._f32(valid_faces)
        out = np.empty(cc.shape, dtype=np.float64)
        self.fn("build_boundary_weights", None)(ccp, lsp, vfp, lp, _res(l), int(axis), out.ctypes.data_as(_f64p))
        return out

    def build_rhs(self, material, velocity, cut_cell, exp_shape, offset, solid_velocity=None):
        m, mp = _i32(material)
        kv, vp = self._ptrs3(velocity)
        kc, cp = self._ptrs3(cut_cell)
        sp = None
        if solid_velocity is not None:
            ks, sp = self._ptrs3(solid_velocity)
        rhs = np.zeros(exp_shape, dtype=np.float64)
        self.fn("build_rhs", None)(mp, vp, cp, sp, _res(m), _res_t(tuple(exp_shape)[::-1]), _res_t(offset), rhs.ctypes.data_as(_f64p))
        return rhs

    def apply_old_pressure(self, pressure, material, exp_shape, offset):
        m, mp = _i32(material)
        p, pp = self._f32(pressure)
        x = np.zeros(exp_shape, dtype=np.float64)
        self.fn("apply_old_pressure", None)(pp, mp, _res(m), _res_t(tuple(exp_shape)[::-1]), _res_t(offset), x.ctypes.data_as(_f64p))
        return x

    def apply_solution_to_pressure(self, pressure, material, solution, offset):
        m, mp = _i32(material)
        p, pp = self._f32(np.array(pressure, dtype=np.float32, copy=True))
        x, xp = _f64(solution)
        self.fn("apply_solution_to_pressure", None)(pp, mp, xp, _res(m), _res(x), _res_t(offset))
        return p

    def apply_pressure_gradient(self, velocity, liquid_surface, pressure, valid_faces, material, axis):
        m, mp = _i32(material)
        v, vp = self._f32(np.array(velocity, dtype=np.float32, copy=True))
        ls, lsp = self._f32(liquid_surface)
        p, pp = self._f32(pressure)
        vf, vfp = self._f32(valid_faces)
        self.fn("apply_pressure_gradient", None)(vp, None, lsp, pp, vfp, mp, _res(m), int(axis))
        return v

    def threads(self):
        return int(self.fn("threads")())

    def set_threads(self, n=None):
        """Use n (default: every host core) OpenMP threads, whatever OMP_NUM_THREADS the launcher exported."""
        self.fn("set_threads", None)(int(n or os.cpu_count() or 1))
        return self.threads()

    def expand_dims(self, base_shape):
        ores = (C.c_int64 * 3)()
        off = (C.c_int64 * 3)()
        lv = C.c_int()
        self.fn("expand_dims", None)(_res_t(base_shape[::-1]), ores, off, C.byref(lv))
        return (ores[2], ores[1], ores[0]), np.array(list(off), dtype=np.int64), int(lv.value)

    def expand_labels(self, base):
        b, bp = _i32(base)
        shape, off, lv = self.expand_dims(b.shape)
        out = np.empty(shape, dtype=np.int32)
        self.fn("expand_labels", None)(bp, _res(b), out.ctypes.data_as(_i32p), _res(out), _res_t(off))
        return out, off, lv

    def expand_weights(self, base_w, base_shape, exp_labels, offset):
        outs = []
        for axis in range(3):
            bw, bwp = _f64(base_w[axis])
            out = np.empty(face_shape(exp_labels.shape, axis), dtype=np.float64)
            self.fn("expand_weights", None)(bwp, _res_t(base_shape[::-1]), _res(exp_labels), _res_t(offset), axis, out.ctypes.data_as(_f64p))
            outs.append(out)
        return outs

    def boundary_cells(self, labels, width=3):
        l, lp = _i32(labels)
        f = self.fn("boundary_cells", C.c_int64)
        n = f(lp, _res(l), int(width), None, C.c_int64(0))
        out = np.empty((max(n, 1), 3), dtype=np.int64)
        f(lp, _res(l), int(width), out.ctypes.data_as(_i64p), C.c_int64(n))
        return out[:n]

    def solver(self, labels, w, levels, use_gs=False, coarse_scale=1.0):
        return _PortSolver(self, labels, w, levels, use_gs, coarse_scale)


class _PortSolver:
    def __init__(self, lib, labels, w, levels, use_gs, coarse_scale):
        self.lib = lib
        self.labels, lp = _i32(labels)
        keep, wp = _wptrs(w)
        f = lib.fn("solver_create", C.c_void_p)
        self.h = C.c_void_p(f(lp, _res(self.labels), *wp, int(levels), int(use_gs), C.c_double(coarse_scale)))
        if not self.h:
            raise RuntimeError("orc_solver_create failed (no solvable level)")

    @property
    def levels(self):
        return int(self.lib.fn("solver_levels")(self.h))

    def level_labels(self, level):
        r = (C.c_int64 * 3)()
        self.lib.fn("solver_level_res", None)(self.h, int(level), r)
        out = np.empty((r[2], r[1], r[0]), dtype=np.int32)
        self.lib.fn("solver_get_labels", None)(self.h, int(level), out.ctypes.data_as(_i32p))
        return out

    def level_boundary_cells(self, level):
        n = self.lib.fn("solver_boundary_count", C.c_int64)(self.h, int(level))
        out = np.empty((max(n, 1), 3), dtype=np.int64)
        self.lib.fn("solver_get_boundary_cells", None)(self.h, int(level), out.ctypes.data_as(_i64p))
        return out[:n]

    @property
    def coarse_unknowns(self):
        return int(self.lib.fn("solver_coarse_unknowns", C.c_int64)(self.h))

    def vcycle(self, x, b, use_initial_guess=False):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        self.lib.fn("solver_vcycle", None)(self.h, xp, bp, int(use_initial_guess))
        return x

    def pcg(self, x, b, tol, max_it, diagonal=False):
        x, xp = _f64(np.array(x, copy=True))
        b, bp = _f64(b)
        hist = np.zeros(max_it + 2, dtype=np.float64)
        cnt = C.c_int()
        it = self.lib.fn("pcg_diag" if diagonal else "pcg")(self.h, xp, bp, C.c_double(tol), int(max_it), hist.ctypes.data_as(_f64p), len(hist),
                                                            C.byref(cnt))
        return x, int(it), hist[: cnt.value].copy()

    def close(self):
        if self.h:
            self.lib.fn("solver_destroy", None)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TestNodeLib:
    """The reference's own diagnostic node (HDK_TestGeometricMultigrid.cpp, compiled unmodified): options in, its log out."""

    __test__ = False  # not a pytest class

    def __init__(self):
        self.lib = C.CDLL(REF_TESTNODE_SO)
        self.lib.ref_testnode_set_threads(int(os.cpu_count() or 1))

    def run(self, **options):
        """options: the node's DOP parameters (HDK_TestGeometricMultigrid.h:10-35), e.g. gridSize=32, useComplexDomain=1, testSymmetry=1."""
        text = ";".join(f"{k}={float(v)!r}" for k, v in options.items())
        buf = C.create_string_buffer(1 << 20)
        ok = self.lib.ref_testnode_run(text.encode(), buf, len(buf))
        return bool(ok), buf.value.decode(errors="replace")

