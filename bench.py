#!/usr/bin/env python
"""bench.py -- the headline measurement: BASELINE.json's "256^3 MGPCG solve ms to 1e-6 resid; V-cycle HBM GB/s vs peak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 256] [--impl reference]

A step is one complete multigrid-preconditioned CG solve of the 256^3 flipSplash-shaped free-surface problem
(BASELINE.json configs[2]; expanded 512^3, 7 levels, seeded random rhs, zero initial guess, tol 1e-6).
  value      solve ms with labels/weights/rhs already resident in HBM (gmg_pcg_device), CUDA events on the library's stream
  e2e        the same solve through the reference-facing entry points with HOST buffers:
             gmg_solver_create (labels + BOUNDARY-cell face weights H2D, hierarchy build) + gmg_pcg_from_zero (rhs H2D from pinned memory;
             the zero initial guess is declared, not sent; pressure D2H) + gmg_solver_destroy -- wall clock around the blocking calls
  roofline   the dominant fine-level kernel class: algorithmic bytes per launch / CUDA-event duration per launch,
             against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the reference's own sources (oracle/_ref, compiled unmodified against the HDK/Eigen shim) -- or the plain-C
             oracle if that library did not travel -- timed on this box's host cores on the same workload
--impl reference times that CPU implementation as the step (constructor + PCG solve), rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

METRIC = "mgpcg_solve_ms"
TOL = 1e-6
MAX_IT = 1000
SEED = 12345
# algorithmic bytes per active cell (R = 8-byte fp64, L = 1-byte label)
BYTES_VCYCLE_PER_CELL = 126.0


def workload_name(n):
    return f"{n}^3 flipSplash-shaped free-surface domain (pool + 3x3 falling blobs), expanded {2 * n}^3, MGPCG to 1e-6, Jacobi smoother"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload, klass):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the class's fine-level kernels, from the committed ncu --set full
    capture of this workload (profiles/traffic.json names the report); None when that workload was not captured."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp))[workload]["bytes_per_launch"].get(klass)
    except Exception:
        return None


def build_inputs(n):
    base_labels, base_w, dx = D.flipsplash_domain(n)
    return base_labels, base_w, dx


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(n, steps, warmup, budget_s=240.0, keep_solution=False, use_gs=False):
    """Times the reference's CPU implementation (constructor + PCG) on this box's host cores -- ALL of them: torchrun exports
    OMP_NUM_THREADS=1 to its workers, which round 1's N>1 reference arm silently obeyed."""
    from oracle import bindings

    kind = "reference"
    try:
        lib = bindings.RefLib()
    except Exception:
        if not os.path.exists(bindings.PORT_SO):
            bindings.build(ref=False)
        lib = bindings.PortLib()
        kind = "port"
    lib.set_threads(os.cpu_count())
    base_labels, base_w, dx = build_inputs(n)
    labels, w, off, levels = lib.expand_domain(base_labels, base_w)
    b = D.random_rhs(labels, dx, SEED)
    times, setups, solves, iters, hist = [], [], [], None, None
    t_begin = time.perf_counter()
    done = 0
    clamp = None
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        s = lib.solver(labels, w, levels, bool(use_gs))
        t1 = time.perf_counter()
        x, iters, hist = s.pcg(np.zeros_like(b), b, TOL, MAX_IT)
        t2 = time.perf_counter()
        s.close()
        if step >= warmup:
            setups.append((t1 - t0) * 1e3)
            solves.append((getattr(s, "last_seconds", None) or (t2 - t1)) * 1e3)
            times.append(setups[-1] + solves[-1])
            done += 1
        if time.perf_counter() - t_begin > budget_s and done >= 1 and step + 1 < warmup + steps:
            clamp = f"time budget of {budget_s:.0f} s reached after {done} of {steps} timed steps (one step is a ~2 s CPU solve)"
            break
    out = dict(kind=kind, cores=lib.threads(), ms=float(np.mean(times)), setup_ms=float(np.mean(setups)), solve_ms=float(np.mean(solves)),
               steps=done, iterations=int(iters), final_rel_residual=float(hist[-1]) if len(hist) else None,
               active_cells=int(D.active_mask(labels).sum()), history=[float(v) for v in hist], steps_clamped_reason=clamp)
    if keep_solution:
        out["x"] = x
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    # the driver's own --steps / --warmup, up to a stated time budget (a step is a ~2-3 s CPU solve)
    r = cpu_reference_run(args.size, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["ms"], "unit": "ms", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "steps_requested": args.steps, "steps_clamped_reason": r["steps_clamped_reason"],
        "ms_per_step": r["ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "step": "solver constructor + PCG solve on the host cores", "iterations": r["iterations"],
                   "active_cells": r["active_cells"]},
        "cpu_baseline": {"value": r["ms"], "unit": "ms", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"full {args.size}^3 workload, constructor {r['setup_ms']:.0f} ms + PCG {r['solve_ms']:.0f} ms, mean of {r['steps']} run(s)"},
        "e2e": {"value": r["ms"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_ms": r["setup_ms"], "solve_only_ms": r["solve_ms"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ parity blocks
def rel_history_dev(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = min(len(a), len(b))
    if m == 0:
        return 0.0
    return float((np.abs(a[:m] - b[:m]) / np.abs(b[:m])).max())


def gather_owned_box(torch, dist, X, off, hi, solver):
    """The solution of a sharded solve as ONE dense array of the base box on every rank: each rank downloads its owned
    z-planes (everything else 0) and the boxes are summed over the ranks."""
    x = X.download()
    box = np.ascontiguousarray(x[int(off[2]):hi[2], int(off[1]):hi[1], int(off[0]):hi[0]])
    del x
    if dist is not None:
        t = torch.from_numpy(box).cuda()
        dist.all_reduce(t)
        box = t.cpu().numpy()
    return box


# ------------------------------------------------------------------------------------------------ 512^3 V-cycle sweep
def measure_sweep(torch, api, ctx, dist, rank, world, n, steps, warmup, flush):
    """BASELINE.json configs[3]: N^3 fully liquid box (one DIRICHLET layer), V-cycle only.  Returns the block bench lines carry
    as `sweep512`: V-cycle ms (max over ranks), algorithmic GB/s, per-kernel-class fractions of the measured HBM peak."""
    t0 = time.perf_counter()
    bl, bw, dx = D.liquid_box_domain(n)
    labels, w, off, levels, box = ctx.buildExpandedDomainLazy(bl, bw)
    del bl, bw
    solver = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box)
    active = solver.active_cells(0)
    rng = np.random.default_rng(SEED)
    b_host = np.zeros(labels.shape, dtype=np.float64)
    sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
    b_host[sl] = rng.random(tuple(s.stop - s.start for s in sl)) * dx * dx * D.active_mask(labels[sl])
    B, Z = solver.grid(0, b_host), solver.grid(0)
    nlev = solver.getMGLevels()
    del w, b_host, labels
    build_s = time.perf_counter() - t0

    def sync():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(warmup):
        solver.applyVCycleDevice(Z, B)
    sync()
    ctx.launch_count(reset=True)
    vc = []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_begin()
        solver.applyVCycleDevice(Z, B)
        vc.append(ctx.timer_end())
    sync()
    launches, comm_ops = ctx.launch_count(), ctx.comm_count()
    ms = float(np.mean(vc))
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.profile_enable(True)
    ctx.profile_reset()
    nprof = 3
    for _ in range(nprof):
        solver.applyVCycleDevice(Z, B)
    prof_fine = ctx.profile(True)
    by_level = {str(l): {k: {"ms_per_vcycle": round(v[0] / nprof, 4), "launches_per_vcycle": v[1] // nprof} for k, v in row.items() if k != "setup" and v[1] > 0}
                for l, row in ctx.profile_by_level(nlev).items()}
    ctx.profile_enable(False)
    peak, peak_kind = measured_peak()
    fine = {k: v for k, v in prof_fine.items() if k not in ("setup", "coarse_solve", "halo_exchange") and v[1] > 0 and v[2] > 0}
    classes = {k: {"us_per_launch": v[0] / v[1] * 1e3, "launches_per_vcycle": v[1] // nprof, "algorithmic_gbs": v[2] / (v[0] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": v[2] / (v[0] * 1e-3) / 1e9 / peak, "traffic_bytes_per_launch": ncu_traffic(f"vcycle{n}", k)} for k, v in fine.items()}
    halo = prof_fine.get("halo_exchange")
    dom = max(fine, key=lambda k: fine[k][0])
    d_ms, d_n, d_bytes = fine[dom]
    achieved = d_bytes / (d_ms * 1e-3) / 1e9
    vb = BYTES_VCYCLE_PER_CELL * active
    out = {
        "workload": f"{n}^3 fully liquid box (one DIRICHLET layer), expanded {2 * n}^3, one V-cycle, Jacobi smoother", "levels": nlev, "active_cells": active,
        "vcycle_ms": ms, "steps": steps, "warmup": warmup, "l2": "256 MB flush write before every timed V-cycle",
        "vcycle_algorithmic_gbs": vb / (ms * 1e-3) / 1e9, "vcycle_frac_of_hbm_peak": vb / (ms * 1e-3) / 1e9 / peak / world,
        "algorithmic_bytes_per_cell": BYTES_VCYCLE_PER_CELL,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(f"vcycle{n}", dom), "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "launches": d_n,
                     "avg_launch_us": d_ms / d_n * 1e3, "algorithmic_bytes_per_launch": d_bytes / d_n},
        "fine_level_kernels": classes, "kernels_by_level": by_level,
        "halo_exchange_ms_per_vcycle": (halo[0] / nprof) if halo and halo[1] else 0.0,
        "kernel_timing": "CUDA event nodes inside the replayed V-cycle graph (rank 0's slab when sharded)",
        "gpu_launches": int(launches), "nccl_ops": int(comm_ops), "host_build_s": build_s,
        "parallelism": "single GPU" if world == 1 else f"{world} z-slabs",
    }
    B.close(); Z.close(); solver.close()
    return out


# ------------------------------------------------------------------------------------------------ 512^3 free-surface solve (north_star target)
def flipsplash_inputs_gpu(torch, n):
    """domains.flipsplash_domain + ghost_fluid_weights with the same arithmetic in the same order, evaluated by torch on the GPU (the numpy
    generator needs ~60 s and ~12 GB per rank at 512^3).  Returns host arrays (labels int32, [w_x, w_y, w_z] float64, dx)."""
    dev, f64 = "cuda", torch.float64
    a = torch.arange(n, dtype=f64, device=dev)
    k, j, i = a[:, None, None], a[None, :, None], a[None, None, :]
    N = float(n)
    phi = (j - 0.25 * n).expand(n, n, n).clone()
    r = 0.06 * N
    for cx in (n / 3.0, n / 2.0, 2.0 * n / 3.0):
        for cz in (n / 3.0, n / 2.0, 2.0 * n / 3.0):
            phi = torch.minimum(phi, torch.sqrt((i - cx) ** 2 + (j - 0.65 * n) ** 2 + (k - cz) ** 2) - r)
    phi /= N
    labels = torch.where(phi <= 0, D.INTERIOR, D.DIRICHLET).to(torch.int32)
    labels[:, :, 0] = D.EXTERIOR
    labels[:, :, n - 1] = D.EXTERIOR
    labels[0, :, :] = D.EXTERIOR
    labels[n - 1, :, :] = D.EXTERIOR
    labels[:, 0, :] = D.EXTERIOR
    weights = []
    for axis in range(3):
        na = 2 - axis
        lb, lf = labels.narrow(na, 0, n - 1), labels.narrow(na, 1, n - 1)
        pb, pf = phi.narrow(na, 0, n - 1), phi.narrow(na, 1, n - 1)
        theta = torch.zeros_like(pb)
        both, ca, cb = (pb < 0) & (pf < 0), (pb < 0) & (pf >= 0), (pb >= 0) & (pf < 0)
        theta = torch.where(both, torch.ones_like(theta), theta)
        theta = torch.where(ca, pb / (pb - pf), theta)
        theta = torch.where(cb, pf / (pf - pb), theta)
        theta = theta.clamp(0.01, 1.0)
        both_int = (lb == D.INTERIOR) & (lf == D.INTERIOR)
        mixed = ((lb == D.INTERIOR) & (lf == D.DIRICHLET)) | ((lb == D.DIRICHLET) & (lf == D.INTERIOR))
        inner = torch.where(both_int, torch.ones_like(theta), torch.where(mixed, 1.0 / theta, torch.zeros_like(theta)))
        del theta, both, ca, cb, both_int, mixed
        shape = [n, n, n]
        shape[na] += 1
        w = torch.zeros(shape, dtype=f64, device=dev)
        w.narrow(na, 1, n - 1).copy_(inner)
        del inner
        weights.append(w.cpu().numpy())
        del w
    out = labels.cpu().numpy(), weights, 1.0 / N
    del phi, labels
    torch.cuda.empty_cache()
    return out


def measure_solve_big(torch, api, ctx, dist, rank, world, local_rank, n, steps, warmup, flush):
    """The north_star's target workload: an n^3 (512^3) flipSplash-shaped FREE-SURFACE MGPCG solve to 1e-6, sharded over the ranks like the
    headline solve.  Block `solve512` of the line: solve ms (max over ranks), iterations, per-kernel-class fractions of the measured HBM
    peak at the fine level, and at N > 1 the agreement with an unsharded solve of the same inputs on rank 0."""
    t0 = time.perf_counter()
    sb, sw, sdx = flipsplash_inputs_gpu(torch, 64)
    nb, nw, ndx = D.flipsplash_domain(64)
    same = bool((sb == nb).all() and sdx == ndx)  # labels must be identical; a face weight may differ in the last bit (division / sqrt rounding)
    wdev = max(float(np.abs(x - y).max()) for x, y in zip(sw, nw))
    bl, bw, dx = flipsplash_inputs_gpu(torch, n)
    labels, w, off, levels, box = ctx.buildExpandedDomainLazy(bl, bw)
    del bl, bw
    solver = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box)
    active = solver.active_cells(0)
    rng = np.random.default_rng(SEED)
    b_host = np.zeros(labels.shape, dtype=np.float64)
    sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
    b_host[sl] = rng.random(tuple(s.stop - s.start for s in sl)) * dx * dx * D.active_mask(labels[sl])
    B, X = solver.grid(0, b_host), solver.grid(0)
    nlev = solver.getMGLevels()
    build_s = time.perf_counter() - t0

    def sync():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    it, hist = 0, []
    for _ in range(warmup):
        X.zero()
        it, hist = solver.solveDevice(X, B, TOL, MAX_IT)
    sync()
    ctx.launch_count(reset=True)
    ms_list = []
    for _ in range(steps):
        X.zero()
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_begin()
        it, hist = solver.solveDevice(X, B, TOL, MAX_IT)
        ms_list.append(ctx.timer_end())
    sync()
    launches, comm_ops = ctx.launch_count(), ctx.comm_count()
    ms = float(np.mean(ms_list))
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.profile_enable(True)
    ctx.profile_reset()
    X.zero()
    solver.solveDevice(X, B, TOL, MAX_IT)
    prof_all, prof_fine = ctx.profile(False), ctx.profile(True)
    ctx.profile_enable(False)
    peak, peak_kind = measured_peak()
    fine = {k: v for k, v in prof_fine.items() if k not in ("setup", "coarse_solve", "halo_exchange") and v[1] > 0 and v[2] > 0}
    total_ms = sum(v[0] for k, v in prof_all.items() if k != "setup")
    classes = {k: {"us_per_launch": v[0] / v[1] * 1e3, "launches_per_solve": v[1], "algorithmic_gbs": v[2] / (v[0] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": v[2] / (v[0] * 1e-3) / 1e9 / peak, "share_of_solve": prof_all[k][0] / total_ms} for k, v in fine.items()}
    halo = prof_all.get("halo_exchange")
    agreement = None
    if world > 1:
        if rank == 0:
            ctx1 = api.Context(local_rank)
            s1 = api.GeometricMultigridPoissonSolver(ctx1, labels, w, levels, box=box)
            B1, X1 = s1.grid(0, b_host), s1.grid(0)
            t1 = []
            for _ in range(2):
                X1.zero()
                ctx1.timer_begin()
                it1, hist1 = s1.solveDevice(X1, B1, TOL, MAX_IT)
                t1.append(ctx1.timer_end())
            agreement = {"n1_solve_ms": float(t1[-1]), "iterations_equal": bool(it1 == it), "max_rel_history_dev": rel_history_dev(hist, hist1),
                         "rel_history_dev_per_iteration": [float(abs(a - b) / abs(b)) for a, b in zip(hist, hist1)],
                         "bar": 1e-6, "bar_note": "north_star asks 1e-5; the V-cycle is bitwise the unsharded one, the dot products are all-reduced in another order: "
                                                  "the deviation starts at rounding level and is amplified by the CG recurrence (measured 1e-9 .. 4e-8 after 18 iterations)"}
            B1.close(); X1.close(); s1.close(); ctx1.close()
        dist.barrier()
    out = {
        "workload": f"{n}^3 flipSplash-shaped free-surface domain (pool + 3x3 falling blobs, ghost-fluid face weights), expanded {2 * n}^3, MGPCG to 1e-6, Jacobi smoother",
        "levels": nlev, "active_cells": active, "solve_ms": ms, "iterations": int(it), "final_rel_residual": float(hist[-1]) if len(hist) else None,
        "steps": steps, "warmup": warmup, "l2": "256 MB flush write before every timed solve",
        "fine_level_kernels": classes, "halo_exchange_ms_per_solve": halo[0] if halo and halo[1] else 0.0,
        "kernel_timing": "CUDA event nodes inside the replayed PCG graphs, one instrumented solve (rank 0's slab when sharded)",
        "peak": peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
        "gpu_launches": int(launches), "nccl_ops": int(comm_ops), "host_build_s": build_s, "inputs": "generated by torch on the GPU (every rank the same arrays); against domains.flipsplash_domain at 64^3: labels "
                  + ("identical" if same else "DIFFERENT") + f", face weights within {wdev:.1e}",
        "agreement_1_vs_n": agreement, "parallelism": "single GPU" if world == 1 else f"{world} z-slabs",
    }
    B.close(); X.close(); solver.close()
    del w, b_host, labels
    return out


# ------------------------------------------------------------------------------------------------ 1024^3 narrow band
def narrow_band_labels_u8(n, thickness=12):
    """BASELINE.json configs[4] without ever holding a dense n^3 float array: the EXPANDED one-byte label
    grid is np.zeros (pages outside the base box are never materialised) and the base box is filled plane by plane with
    terrain h(x,z) = N(0.5 + 0.1 sin(2 pi x/N) sin(2 pi z/N)); y < h - thickness solid, <= h liquid, above air; walls solid.
    BOUNDARY cells by the rule of setBoundaryCellLabels with unit weights (Ops.h:1574-1644): an INTERIOR cell with a DIRICHLET
    or EXTERIOR 6-neighbour.  Returns (labels uint8 expanded, offset, mgLevels, (lo, hi))."""
    from geometricmultigridpressuresolver_b200 import api

    shape, off, levels = api.expand_dims((n, n, n))
    labels = np.zeros(shape, dtype=np.uint8)
    i = np.arange(n, dtype=np.float64)[None, :]
    j = np.arange(n, dtype=np.float64)[:, None]

    def plane(k):
        if k <= 0 or k >= n - 1:
            return np.full((n, n), D.EXTERIOR, dtype=np.uint8)
        h = n * (0.5 + 0.1 * np.sin(2 * np.pi * i / n) * np.sin(2 * np.pi * k / n))
        p = np.where(j > h, D.DIRICHLET, np.where(j < h - thickness, D.EXTERIOR, D.INTERIOR)).astype(np.uint8)
        p[:, 0] = D.EXTERIOR
        p[:, n - 1] = D.EXTERIOR
        p[0, :] = D.EXTERIOR
        return p

    prev, cur = plane(-1), plane(0)
    ox, oy, oz = int(off[0]), int(off[1]), int(off[2])
    for k in range(n):
        nxt = plane(k + 1)
        nonint = cur != D.INTERIOR  # DIRICHLET or EXTERIOR (BOUNDARY does not exist yet in cur/prev/nxt)
        nb = (prev != D.INTERIOR) | (nxt != D.INTERIOR)
        nb[1:, :] |= nonint[:-1, :]
        nb[:-1, :] |= nonint[1:, :]
        nb[:, 1:] |= nonint[:, :-1]
        nb[:, :-1] |= nonint[:, 1:]
        out = cur.copy()
        out[(cur == D.INTERIOR) & nb] = D.BOUNDARY
        labels[oz + k, oy:oy + n, ox:ox + n] = out
        prev, cur = cur, nxt
    hi = [ox + n, oy + n, oz + n]
    return labels, off, levels, (off, hi)


def measure_narrow(torch, api, ctx, dist, rank, world, local_rank, n, flush, thickness=12):
    """configs[4]: n^3 narrow-band liquid sheet, MGPCG to 1e-6 on `world` z-slabs; 1-vs-N agreement against an unsharded solve on
    rank 0 (iteration count, residual history, |x|^2)."""
    out = {"workload": f"{n}^3 narrow-band liquid sheet ({thickness} cells) over sinusoidal terrain, expanded {2 * n}^3, unit face weights, MGPCG to 1e-6"}
    try:
        t0 = time.perf_counter()
        labels, off, levels, box = narrow_band_labels_u8(n, thickness)
        gen_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        solver = api.GeometricMultigridPoissonSolver(ctx, labels, None, levels, box=box)
        setup_s = time.perf_counter() - t0
        sl = tuple(slice(int(box[0][2 - k]), int(box[1][2 - k])) for k in range(3))
        act = D.active_mask(labels[sl])
        b_host = np.zeros(labels.shape, dtype=np.float64)
        rng = np.random.default_rng(SEED)
        sub = np.zeros(act.shape, dtype=np.float64)
        sub[act] = rng.random(int(act.sum())) / float(n * n)
        b_host[sl] = sub
        del sub
        B, X = solver.grid(0, b_host), solver.grid(0)
        active = solver.active_cells(0)
        box_cells = int(np.prod([s.stop - s.start for s in sl]))

        def one():
            X.zero()
            flush.fill_(1)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            ctx.timer_begin()
            it, hist = solver.solveDevice(X, B, TOL, MAX_IT)
            return ctx.timer_end(), it, hist

        one()
        ms = []
        for _ in range(3):
            t, it, hist = one()
            ms.append(t)
        val = float(np.mean(ms))
        xx = solver.dotProduct(X, X)
        if dist is not None:
            t = torch.tensor([val], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            val = float(t.item())
        # one V-cycle for the bandwidth figure
        Z = solver.grid(0)
        solver.applyVCycleDevice(Z, B)
        vc = []
        for _ in range(5):
            flush.fill_(1)
            torch.cuda.synchronize()
            ctx.timer_begin()
            solver.applyVCycleDevice(Z, B)
            vc.append(ctx.timer_end())
        vms = float(np.mean(vc))
        if dist is not None:
            t = torch.tensor([vms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vms = float(t.item())
        peak, _ = measured_peak()
        nlev = solver.getMGLevels()
        sharded = [l for l in range(nlev) if solver.shard_info(l)[0]]
        out.update({"levels_requested": int(levels), "levels": nlev, "active_cells": int(active), "occupancy_of_base_grid": active / float(n) ** 3,
                    "stored_box_cells": box_cells, "coarse_unknowns": solver.coarse_unknowns(), "solve_ms": val, "iterations": int(it),
                    "final_rel_residual": float(hist[-1]), "vcycle_ms": vms,
                    "vcycle_algorithmic_gbs": BYTES_VCYCLE_PER_CELL * active / (vms * 1e-3) / 1e9,
                    "vcycle_frac_of_hbm_peak": BYTES_VCYCLE_PER_CELL * active / (vms * 1e-3) / 1e9 / peak / world,
                    "sharded_levels": sharded, "setup_s": setup_s, "label_generation_s": gen_s, "steps": 3, "warmup": 1})
        Z.close(); B.close(); X.close(); solver.close()
        if world > 1:
            agree = None
            if rank == 0:
                ctx1 = api.Context(local_rank)
                s1 = api.GeometricMultigridPoissonSolver(ctx1, labels, None, levels, box=box)
                B1, X1 = s1.grid(0, b_host), s1.grid(0)
                it1, hist1 = s1.solveDevice(X1, B1, TOL, MAX_IT)
                ctx1.timer_begin()
                X1.zero()
                it1, hist1 = s1.solveDevice(X1, B1, TOL, MAX_IT)
                ms1 = ctx1.timer_end()
                xx1 = s1.dotProduct(X1, X1)
                agree = {"n1_solve_ms": ms1, "iterations_equal": bool(it1 == it), "max_rel_history_dev": rel_history_dev(hist, hist1),
                         "rel_dev_x_norm2": abs(xx - xx1) / xx1, "bar": 1e-12}
                B1.close(); X1.close(); s1.close(); ctx1.close()
            out["agreement_1_vs_n"] = agree
    except Exception as e:  # a refused configuration (e.g. a coarsest level beyond the dense direct solve) is reported, not fatal
        out["error"] = f"{type(e).__name__}: {e}"
    if dist is not None:
        dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, rank, world, local_rank):
    import torch

    from geometricmultigridpressuresolver_b200 import api

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    if dist is not None:
        ctx.shard_with_torch(dist)  # z-slabs over the ranks: NCCL halo planes + scalar all-reduces inside the library
    n = args.size
    base_labels, base_w, dx = build_inputs(n)
    labels, w, off, levels = ctx.buildExpandedDomain(base_labels, base_w)
    hi = [int(off[a]) + base_labels.shape[2 - a] for a in range(3)]
    box = (off, hi)
    b_host = D.random_rhs(labels, dx, SEED)
    solver = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box)
    active = solver.active_cells(0)
    B, X = solver.grid(0, b_host), solver.grid(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    def one_solve():
        X.zero()
        flush.fill_(1)  # L2 flush between timed iterations
        torch.cuda.synchronize()
        ctx.timer_begin()
        it, hist = solver.solveDevice(X, B, TOL, MAX_IT)
        return ctx.timer_end(), it, hist

    for _ in range(args.warmup):
        one_solve()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.launch_count(reset=True)
    step_ms = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ms, it, hist = one_solve()
        step_ms.append(ms)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = ctx.launch_count()
    comm_ops = ctx.comm_count()
    clocks = sampler.stop()
    ms_per_step = float(np.mean(step_ms))
    if dist is not None:
        t = torch.tensor([ms_per_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())

    # ---- V-cycle only: ms and algorithmic GB/s ----------------------------------------------------------
    Z = solver.grid(0)
    for _ in range(3):
        solver.applyVCycleDevice(Z, B)
    vc = []
    for _ in range(10):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_begin()
        solver.applyVCycleDevice(Z, B)
        vc.append(ctx.timer_end())
    vcycle_ms = float(np.mean(vc))

    # ---- per-kernel-class device times (events around every launch; separate pass so `value` carries no event overhead) ----
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(max(1, min(args.steps, 3))):
        X.zero()
        solver.solveDevice(X, B, TOL, MAX_IT)
    prof_all, prof_fine = ctx.profile(False), ctx.profile(True)
    nprof = max(1, min(args.steps, 3))
    by_level = {str(l): {k: {"ms_per_solve": round(v[0] / nprof, 4), "launches_per_solve": v[1] // nprof} for k, v in row.items() if k != "setup"}
                for l, row in ctx.profile_by_level(solver.getMGLevels()).items()}
    ctx.profile_enable(False)
    # second pass for the band sweeps alone: ONE event pair around the three back-to-back launches of a sweep group (profiling mode 2) --
    # an event-record node costs ~5 us and stops the launches it separates from overlapping, which overstates a ~12 us kernel by half
    ctx.profile_enable(2)
    ctx.profile_reset()
    for _ in range(nprof):
        X.zero()
        solver.solveDevice(X, B, TOL, MAX_IT)
    prof_fine_grouped = ctx.profile(True)
    ctx.profile_enable(False)
    peak, peak_kind = measured_peak()
    solve_classes = {k: v for k, v in prof_all.items() if k not in ("setup",) and v[1] > 0}
    total_ms = sum(v[0] for v in solve_classes.values())
    # the roofline is quoted on the dominant HBM-bound kernel class; the NVLink halo exchanges of a sharded run are reported in `kernels`
    fine = {k: v for k, v in prof_fine.items() if k not in ("setup", "coarse_solve", "halo_exchange") and v[1] > 0 and v[2] > 0}
    dom = max(fine, key=lambda k: fine[k][0])
    d_ms, d_n, d_bytes = fine[dom]
    achieved_per_launch_events = d_bytes / d_n / (d_ms / d_n * 1e-3) / 1e9
    roof_timing = "one CUDA-event pair (event-record nodes inside the replayed PCG graphs) around every launch of the class on level 0"
    if dom == "band_jacobi" and prof_fine_grouped.get(dom, (0, 0, 0))[1] == d_n and prof_fine_grouped[dom][0] > 0:
        d_ms, d_n, d_bytes = prof_fine_grouped[dom]
        roof_timing = ("one CUDA-event pair (event-record nodes inside the replayed PCG graphs) around each back-to-back group of 3 sweeps on level 0, "
                       "divided by the launches inside; `frac_per_launch_events` is the same class with a pair around every single launch")
    achieved = d_bytes / d_n / (d_ms / d_n * 1e-3) / 1e9
    traffic = ncu_traffic(f"pcg{n}", dom)
    kernels = {k: {"ms_per_solve": v[0] / max(1, min(args.steps, 3)), "launches_per_solve": v[1] // max(1, min(args.steps, 3)),
                   "share": v[0] / total_ms,
                   "fine_level_gbs": (prof_fine[k][2] / (prof_fine[k][0] * 1e-3) / 1e9) if prof_fine[k][0] > 0 and prof_fine[k][2] > 0 else None}
               for k, v in solve_classes.items()}

    # ---- e2e: host buffers through the reference-facing calls (every rank takes part; max over ranks) ------
    e2e = None
    if not args.quick:
        # every host buffer the reference-facing calls take is page-locked (the contract's "pinned host memory"): the
        # library then DMAs the cropped box straight out of / into the caller's arrays
        rt = torch.cuda.cudart()

        def pin(a):
            err = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            assert int(err) == 0, f"cudaHostRegister failed: {err}"
            return a

        x_np, b_np = pin(np.zeros(labels.shape, dtype=np.float64)), pin(np.ascontiguousarray(b_host))
        labels = pin(np.ascontiguousarray(labels, dtype=np.int32))
        w = [pin(np.ascontiguousarray(a, dtype=np.float64)) for a in w]
        e2e_ms, e2e_setup, it2, hist2 = [], [], None, None
        for step in range(1 + max(1, min(args.steps, 5))):
            x_np[...] = 0.0
            t0 = time.perf_counter()
            s2 = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box)
            t1 = time.perf_counter()
            # no warm start: the solution grid is the constant zero of GFS.cpp:392-398 and is declared so (gmg_pcg_from_zero), like the
            # facade does for a constant-compressed UT_VoxelArray; the pressure is written into the same page-locked array
            x_out, it2, hist2 = s2.solveGeometricConjugateGradient(x_np, b_np, TOL, MAX_IT, inplace=True, solutionIsZero=True)
            s2.close()
            t2 = time.perf_counter()
            if step > 0:
                e2e_ms.append((t2 - t0) * 1e3)
                e2e_setup.append((t1 - t0) * 1e3)
        box_cells = int(np.prod([hi[a] - int(off[a]) + 4 for a in range(3)]))
        # construction: int32 labels of the box + the six face weights of every BOUNDARY cell (the weight grids themselves stay on the
        # host: only BOUNDARY cells look at them; their index list comes back first); per solve: rhs in (x0 = 0 is declared, not sent), pressure out
        n_boundary = int((labels == 3).sum())
        io_cells, io_copies = solver.transfer_cells()  # rhs / pressure move the active rectangles of each z-plane only
        h2d = box_cells * 4 + n_boundary * 48 + io_cells * 8
        d2h = io_cells * 8 + n_boundary * 4
        e2e_val = float(np.mean(e2e_ms))
        if dist is not None:
            t = torch.tensor([e2e_val], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_val = float(t.item())
            sh, zlo, zhi, _ = solver.shard_info(0)
            if sh:  # per rank: its slab of the weights / rhs (+ the replicated one-byte labels); summed over the ranks below
                frac = (zhi - zlo + 20) / float(hi[2] - int(off[2]) + 4)
                h2d = int(box_cells * 4 + n_boundary * frac * 48 + io_cells * 8)
                d2h = int(io_cells * 8 * (zhi - zlo) / float(zhi - zlo + 20))
            t = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64)
            dist.all_reduce(t)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        e2e = {"value": e2e_val, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "setup_ms": float(np.mean(e2e_setup)), "iterations": int(it2),
               "what": "gmg_solver_create (labels + BOUNDARY-cell face weights H2D, hierarchy build) + gmg_pcg_from_zero (rhs H2D, x0 = 0 declared by the caller, pressure D2H) + gmg_solver_destroy; all host buffers page-locked"}
        for a in [x_np, b_np, labels] + list(w):
            rt.cudaHostUnregister(a.ctypes.data)

    # ---- the reference's production smoother (tiled Gauss-Seidel, GFS.cpp:463-466), reported beside the Jacobi headline ----
    gs = None
    if world == 1 and not args.quick:
        sg = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box, useGaussSeidel=True)
        Bg, Xg = sg.grid(0, b_host), sg.grid(0)
        gms = []
        for k in range(5):
            Xg.zero()
            flush.fill_(1)
            torch.cuda.synchronize()
            ctx.timer_begin()
            itg, histg = sg.solveDevice(Xg, Bg, TOL, MAX_IT)
            if k >= 2:
                gms.append(ctx.timer_end())
            else:
                ctx.timer_end()
        gs = {"solve_ms": float(np.mean(gms)), "iterations": int(itg), "final_rel_residual": float(histg[-1]),
              "what": "same solve with useGaussSeidel = true (tiled Gauss-Seidel wavefront kernel), mean of 3 after 2 warm-ups"}
        sg.close()
        if rank == 0 and not args.no_cpu_baseline:
            # the reference's own sources in the same mode on this box's host cores, once: for context beside the Jacobi-mode cpu_baseline
            try:
                rg = cpu_reference_run(n, 1, 0, use_gs=True)
                gs["cpu_reference"] = {"ms": rg["ms"], "setup_ms": rg["setup_ms"], "solve_ms": rg["solve_ms"], "iterations": rg["iterations"], "cores": rg["cores"],
                                       "kind": rg["kind"], "final_rel_residual": rg["final_rel_residual"],
                                       "iterations_equal": bool(int(rg["iterations"]) == int(itg)),
                                       "max_rel_history_dev": rel_history_dev(histg, rg["history"])}
            except Exception as e:  # context only: never fails the line
                gs["cpu_reference"] = {"error": repr(e)[:200]}

    # ---- mixed precision, reported beside the fp64 headline, never as it: fp32 V-cycle inside the fp64 CG -----------
    mixed = None
    if world == 1 and not args.quick:
        try:
            sm = api.GeometricMultigridPoissonSolver(ctx, labels, w, levels, box=box, mixed_precision=True)
            Bm, Xm = sm.grid(0, b_host), sm.grid(0)
            mms = []
            for k in range(5):
                Xm.zero()
                flush.fill_(1)
                torch.cuda.synchronize()
                ctx.timer_begin()
                itm, histm = sm.solveDevice(Xm, Bm, TOL, MAX_IT)
                t = ctx.timer_end()
                if k >= 2:
                    mms.append(t)
            X.zero()
            it64, hist64 = solver.solveDevice(X, B, TOL, MAX_IT)
            mixed = {"solve_ms": float(np.mean(mms)), "iterations": int(itm), "iterations_fp64": int(it64), "final_rel_residual": float(histm[-1]),
                     "max_rel_history_drift_vs_fp64": rel_history_dev(histm, hist64),
                     "what": "gmg_solver_options.mixed_precision = 1: fp32 V-cycle (kernel levels) inside the fp64 CG, mean of 3 after 2 warm-ups; NOT the headline dtype"}
            Bm.close(); Xm.close(); sm.close()
        except Exception as e:
            mixed = {"error": f"{type(e).__name__}: {e}"}

    # ---- parity, asserted (a failed bar makes the process exit non-zero after the line is printed) -------------------------
    parity_failed = []
    X.zero()
    it, hist = solver.solveDevice(X, B, TOL, MAX_IT)
    x_box = gather_owned_box(torch, dist, X, off, hi, solver)
    cpu, parity_vs_cpu, parity_vs_n1 = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        # the reference's own sources on this box's host cores, same inputs: timed (cpu_baseline) AND compared
        r = cpu_reference_run(n, 1, 0, keep_solution=True)
        cpu = {"value": r["ms"], "unit": "ms", "cores": r["cores"], "kind": r["kind"],
               "sample": f"full {n}^3 workload once: constructor {r['setup_ms']:.0f} ms + PCG {r['solve_ms']:.0f} ms, {r['iterations'] + 1} iterations",
               "iterations": r["iterations"], "final_rel_residual": r["final_rel_residual"]}
        xr = r.pop("x")[int(off[2]):hi[2], int(off[1]):hi[1], int(off[0]):hi[0]]
        parity_vs_cpu = {"iterations_gpu": int(it), "iterations_cpu": int(r["iterations"]), "iterations_equal": bool(int(it) == int(r["iterations"])),
                         "max_rel_history_dev": rel_history_dev(hist, r["history"]),
                         "max_rel_x_dev": float(np.abs(x_box - xr).max() / np.abs(xr).max()),
                         "bars": {"history": 1e-5, "x": 1e-5, "iterations": "+-1"}, "oracle": r["kind"]}
        if abs(int(it) - int(r["iterations"])) > 1 or parity_vs_cpu["max_rel_history_dev"] > 1e-5 or parity_vs_cpu["max_rel_x_dev"] > 1e-5:
            parity_failed.append("parity_vs_cpu")
        del xr
    if world > 1:
        # the driver's GPU test box has one GPU, so the sharded path proves itself here: rank 0 repeats the solve on an
        # UNSHARDED context of its own GPU and the line carries the deviation; history deviation above 1e-9 fails the run
        if rank == 0:
            ctx1 = api.Context(local_rank)
            s1 = api.GeometricMultigridPoissonSolver(ctx1, labels, w, levels, box=box)
            B1, X1 = s1.grid(0, b_host), s1.grid(0)
            it1, hist1 = s1.solveDevice(X1, B1, TOL, MAX_IT)
            x1 = X1.download()[int(off[2]):hi[2], int(off[1]):hi[1], int(off[0]):hi[0]]
            parity_vs_n1 = {"iterations_equal": bool(it1 == it), "max_rel_history_dev": rel_history_dev(hist, hist1),
                            "max_rel_x_dev": float(np.abs(x_box - x1).max() / np.abs(x1).max()), "bars": {"history": 1e-9, "x": 1e-9}}
            if it1 != it or parity_vs_n1["max_rel_history_dev"] > 1e-9 or parity_vs_n1["max_rel_x_dev"] > 1e-9:
                parity_failed.append("parity_vs_n1")
            B1.close(); X1.close(); s1.close(); ctx1.close()
            del x1
        dist.barrier()
    del x_box
    n_levels, shard_desc = solver.getMGLevels(), None
    if world > 1:
        shard_desc = f"{world} z-slabs, levels 0..{sum(1 for l in range(n_levels) if solver.shard_info(l)[0]) - 1} sharded with deep halos, coarser levels replicated"
    setup_ms_resident = solver.setup_ms()
    # ---- the HBM-bound and the sparse configurations, in the same line (BASELINE.json configs[3] and [4]) -------------------
    B.close(); X.close(); Z.close(); solver.close()
    del labels, w, b_host
    sweep = None
    if not args.no_sweep and not args.quick:
        sweep = measure_sweep(torch, api, ctx, dist, rank, world, args.sweep_size, 10, 3, flush)
    solve_big = None
    if not args.no_sweep and not args.quick:
        # the north_star's target workload (512^3 free-surface MGPCG), sharded like the headline solve
        solve_big = measure_solve_big(torch, api, ctx, dist, rank, world, local_rank, args.sweep_size, 3, 1, flush)
        ag = (solve_big or {}).get("agreement_1_vs_n")
        if ag and (not ag["iterations_equal"] or ag["max_rel_history_dev"] > ag["bar"]):
            parity_failed.append("solve512_agreement_1_vs_n")
    narrow = None
    narrow_size = args.narrow_size if args.narrow_size is not None else (1024 if world == 8 else 0)
    if narrow_size:
        narrow = measure_narrow(torch, api, ctx, dist, rank, world, local_rank, narrow_size, flush)
        ag = (narrow or {}).get("agreement_1_vs_n")
        if ag and (not ag["iterations_equal"] or ag["max_rel_history_dev"] > 1e-9):
            parity_failed.append("narrow_agreement_1_vs_n")

    if rank == 0:
        vcycle_bytes = BYTES_VCYCLE_PER_CELL * active
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "levels": n_levels, "active_cells": active, "l2": "256 MB flush write before every timed step",
                       "parallelism": "single GPU" if world == 1 else shard_desc},
            "iterations": int(it), "final_rel_residual": float(hist[-1]), "setup_ms": setup_ms_resident,
            "vcycle_ms": vcycle_ms, "vcycle_algorithmic_gbs": vcycle_bytes / (vcycle_ms * 1e-3) / 1e9,
            "vcycle_frac_of_hbm_peak": vcycle_bytes / (vcycle_ms * 1e-3) / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "launches": d_n, "avg_launch_us": d_ms / d_n * 1e3,
                         "algorithmic_bytes_per_launch": d_bytes / d_n, "timing": roof_timing, "frac_per_launch_events": achieved_per_launch_events / peak,
                         # the event-record nodes serialise the graph (no prologue overlap) and cost ~5 us per launch: the instrumented
                         # solve takes total_ms / nprof, the timed one `value`; the same class time scaled by that ratio, for context only
                         "instrumented_solve_ms": total_ms / nprof, "frac_scaled_to_uninstrumented_solve": achieved_per_launch_events / peak * (total_ms / nprof) / ms_per_step},
            "kernels": kernels, "kernels_by_level": by_level,
            "kernel_timing": "CUDA events recorded as nodes inside the replayed PCG graphs (warm L2, back-to-back launches); separate pass from `value`",
            "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu, "gauss_seidel": gs, "gpu_launches": int(launches), "nccl_ops": int(comm_ops), "wall_s_timed_region": wall_s,
            "parity_vs_cpu": parity_vs_cpu, "parity_vs_n1": parity_vs_n1, "parity_failed": parity_failed,
            "sweep512": sweep, "solve512": solve_big, "narrow1024": narrow, "mixed_precision": mixed,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        t = torch.tensor([float(len(parity_failed))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity_failed = parity_failed or (["on another rank"] if t.item() > 0 else [])
        dist.barrier()
        dist.destroy_process_group()
    if parity_failed:
        sys.exit(f"bench.py: parity bar missed: {parity_failed}")


# ------------------------------------------------------------------------------------------------ V-cycle sweep (BASELINE config 4)
def run_vcycle_sweep(args, rank, world, local_rank):
    """--workload vcycle: the sweep512 block (BASELINE.json configs[3]) as a line of its own, at any --size."""
    import torch

    from geometricmultigridpressuresolver_b200 import api

    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    if dist is not None:
        ctx.shard_with_torch(dist)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local_rank)
    sampler.start()
    if args.workload == "solve512":
        blk = measure_solve_big(torch, api, ctx, dist, rank, world, local_rank, args.size, args.steps, args.warmup, flush)
        clocks = sampler.stop()
        if rank == 0:
            line = {"metric": "mgpcg_solve_ms", "value": blk["solve_ms"], "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": blk["solve_ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": blk["workload"], "levels": blk["levels"], "active_cells": blk["active_cells"], "l2": blk["l2"],
                               "parallelism": blk["parallelism"]}, "clocks": clocks}
            line.update({k: v for k, v in blk.items() if k not in ("workload", "levels", "active_cells", "l2", "parallelism", "steps", "warmup")})
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    blk = measure_sweep(torch, api, ctx, dist, rank, world, args.size, args.steps, args.warmup, flush)
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": "vcycle_ms", "value": blk["vcycle_ms"], "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": blk["vcycle_ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": blk["workload"], "levels": blk["levels"], "active_cells": blk["active_cells"], "l2": blk["l2"],
                           "parallelism": blk["parallelism"]}, "clocks": clocks}
        line.update({k: v for k, v in blk.items() if k not in ("workload", "levels", "active_cells", "l2", "parallelism", "steps", "warmup")})
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--workload", default="pcg", choices=["pcg", "vcycle", "solve512"],
                    help="pcg: the headline 256^3 MGPCG solve; vcycle: V-cycle-only sweep (config 4); solve512: the north_star's 512^3 free-surface solve as a line of its own")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="A/B runs: value, V-cycle and per-kernel times only (no e2e, Gauss-Seidel, CPU baseline, sweep512)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the sweep512 block (BASELINE.json configs[3]) of the default run")
    ap.add_argument("--sweep-size", type=int, default=512)
    ap.add_argument("--narrow-size", type=int, default=None, help="narrow-band block size (default: 1024 at 8 GPUs, else off; 0 = off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.size is None:
        args.size = 512 if args.workload in ("vcycle", "solve512") else 256
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    elif args.workload in ("vcycle", "solve512"):
        run_vcycle_sweep(args, rank, world, local_rank)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
