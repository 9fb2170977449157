#!/usr/bin/env python
"""bench.py -- the headline measurement: BASELINE.json's "256^3 MGPCG solve ms to 1e-6 resid; V-cycle HBM GB/s vs peak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 256] [--impl reference]

A step is one complete multigrid-preconditioned CG solve of the 256^3 flipSplash-shaped free-surface problem
(BASELINE.json configs[2]; expanded 512^3, 7 levels, seeded random rhs, zero initial guess, tol 1e-6).
  value      solve ms with labels/weights/rhs already resident in HBM (gmg_pcg_device), CUDA events on the library's stream
  e2e        the same solve through the reference-facing entry points with HOST buffers:
             gmg_solver_create (labels + BOUNDARY-cell face weights H2D, hierarchy build) + gmg_pcg (rhs/x0 H2D from pinned memory,
             pressure D2H) + gmg_solver_destroy -- wall clock around the blocking calls
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
def cpu_reference_run(n, steps, warmup, budget_s=240.0):
    """Times the reference's CPU implementation (constructor + PCG) on this box's host cores."""
    from oracle import bindings

    kind = "reference"
    try:
        lib = bindings.RefLib()
    except Exception:
        if not os.path.exists(bindings.PORT_SO):
            bindings.build(ref=False)
        lib = bindings.PortLib()
        kind = "port"
    base_labels, base_w, dx = build_inputs(n)
    labels, w, off, levels = lib.expand_domain(base_labels, base_w)
    b = D.random_rhs(labels, dx, SEED)
    times, setups, solves, iters, hist = [], [], [], None, None
    t_begin = time.perf_counter()
    done = 0
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        s = lib.solver(labels, w, levels, False)
        t1 = time.perf_counter()
        x, iters, hist = s.pcg(np.zeros_like(b), b, TOL, MAX_IT)
        t2 = time.perf_counter()
        s.close()
        if step >= warmup:
            setups.append((t1 - t0) * 1e3)
            solves.append((getattr(s, "last_seconds", None) or (t2 - t1)) * 1e3)
            times.append(setups[-1] + solves[-1])
            done += 1
        if time.perf_counter() - t_begin > budget_s and done >= 1:
            break
    return dict(kind=kind, cores=lib.threads(), ms=float(np.mean(times)), setup_ms=float(np.mean(setups)), solve_ms=float(np.mean(solves)),
                steps=done, iterations=int(iters), final_rel_residual=float(hist[-1]) if len(hist) else None,
                active_cells=int(D.active_mask(labels).sum()))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference_run(args.size, min(args.steps, 3), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["ms"], "unit": "ms", "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "step": "solver constructor + PCG solve on the host cores", "iterations": r["iterations"],
                   "active_cells": r["active_cells"]},
        "cpu_baseline": {"value": r["ms"], "unit": "ms", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"full {args.size}^3 workload, constructor {r['setup_ms']:.0f} ms + PCG {r['solve_ms']:.0f} ms, mean of {r['steps']} run(s)"},
        "e2e": {"value": r["ms"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_ms": r["setup_ms"], "solve_only_ms": r["solve_ms"], "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    peak, peak_kind = measured_peak()
    solve_classes = {k: v for k, v in prof_all.items() if k not in ("setup",) and v[1] > 0}
    total_ms = sum(v[0] for v in solve_classes.values())
    # the roofline is quoted on the dominant HBM-bound kernel class; the NVLink halo exchanges of a sharded run are reported in `kernels`
    fine = {k: v for k, v in prof_fine.items() if k not in ("setup", "coarse_solve", "halo_exchange") and v[1] > 0 and v[2] > 0}
    dom = max(fine, key=lambda k: fine[k][0])
    d_ms, d_n, d_bytes = fine[dom]
    achieved = d_bytes / d_n / (d_ms / d_n * 1e-3) / 1e9
    traffic = ncu_traffic(f"pcg{n}", dom)
    kernels = {k: {"ms_per_solve": v[0] / max(1, min(args.steps, 3)), "launches_per_solve": v[1] // max(1, min(args.steps, 3)),
                   "share": v[0] / total_ms,
                   "fine_level_gbs": (prof_fine[k][2] / (prof_fine[k][0] * 1e-3) / 1e9) if prof_fine[k][0] > 0 and prof_fine[k][2] > 0 else None}
               for k, v in solve_classes.items()}

    # ---- e2e: host buffers through the reference-facing calls (every rank takes part; max over ranks) ------
    e2e = None
    if True:
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
            x_out, it2, hist2 = s2.solveGeometricConjugateGradient(x_np, b_np, TOL, MAX_IT, inplace=True)
            s2.close()
            t2 = time.perf_counter()
            if step > 0:
                e2e_ms.append((t2 - t0) * 1e3)
                e2e_setup.append((t1 - t0) * 1e3)
        box_cells = int(np.prod([hi[a] - int(off[a]) + 4 for a in range(3)]))
        # construction: int32 labels of the box + the six face weights of every BOUNDARY cell (the weight grids themselves stay on the
        # host: only BOUNDARY cells look at them; their index list comes back first); per solve: rhs + x0 in, pressure out
        n_boundary = int((labels == 3).sum())
        io_cells, io_copies = solver.transfer_cells()  # rhs / x0 / pressure move the active rectangles of each z-plane only
        h2d = box_cells * 4 + n_boundary * 48 + 2 * io_cells * 8
        d2h = io_cells * 8 + n_boundary * 4
        e2e_val = float(np.mean(e2e_ms))
        if dist is not None:
            t = torch.tensor([e2e_val], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_val = float(t.item())
            sh, zlo, zhi, _ = solver.shard_info(0)
            if sh:  # per rank: its slab of the weights / rhs / x0 (+ the replicated one-byte labels); summed over the ranks below
                frac = (zhi - zlo + 20) / float(hi[2] - int(off[2]) + 4)
                h2d = int(box_cells * 4 + n_boundary * frac * 48 + io_cells * 2 * 8)
                d2h = int(io_cells * 8 * (zhi - zlo) / float(zhi - zlo + 20))
            t = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64)
            dist.all_reduce(t)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        e2e = {"value": e2e_val, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "setup_ms": float(np.mean(e2e_setup)), "iterations": int(it2),
               "what": "gmg_solver_create (labels + BOUNDARY-cell face weights H2D, hierarchy build) + gmg_pcg (rhs/x0 H2D, pressure D2H) + gmg_solver_destroy; all host buffers page-locked"}
        for a in [x_np, b_np, labels] + list(w):
            rt.cudaHostUnregister(a.ctypes.data)

    # ---- the reference's production smoother (tiled Gauss-Seidel, GFS.cpp:463-466), reported beside the Jacobi headline ----
    gs = None
    if world == 1:
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(n, 1, 0)
        cpu = {"value": r["ms"], "unit": "ms", "cores": r["cores"], "kind": r["kind"],
               "sample": f"full {n}^3 workload once: constructor {r['setup_ms']:.0f} ms + PCG {r['solve_ms']:.0f} ms, {r['iterations'] + 1} iterations",
               "iterations": r["iterations"], "final_rel_residual": r["final_rel_residual"]}

    if rank == 0:
        vcycle_bytes = BYTES_VCYCLE_PER_CELL * active
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "levels": solver.getMGLevels(), "active_cells": active, "l2": "256 MB flush write before every timed step",
                       "parallelism": "single GPU" if world == 1 else
                       f"{world} z-slabs over NCCL, levels 0..{sum(1 for l in range(solver.getMGLevels()) if solver.shard_info(l)[0]) - 1} sharded with deep halos, coarser levels replicated"},
            "iterations": int(it), "final_rel_residual": float(hist[-1]), "setup_ms": solver.setup_ms(),
            "vcycle_ms": vcycle_ms, "vcycle_algorithmic_gbs": vcycle_bytes / (vcycle_ms * 1e-3) / 1e9,
            "vcycle_frac_of_hbm_peak": vcycle_bytes / (vcycle_ms * 1e-3) / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "launches": d_n, "avg_launch_us": d_ms / d_n * 1e3,
                         "algorithmic_bytes_per_launch": d_bytes / d_n},
            "kernels": kernels, "kernels_by_level": by_level,
            "kernel_timing": "CUDA events recorded as nodes inside the replayed PCG graphs (warm L2, back-to-back launches); separate pass from `value`",
            "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu, "gauss_seidel": gs, "gpu_launches": int(launches), "nccl_ops": int(comm_ops), "wall_s_timed_region": wall_s,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ V-cycle sweep (BASELINE config 4)
def run_vcycle_sweep(args, rank, world, local_rank):
    """--workload vcycle: N^3 fully liquid box (BASELINE.json configs[3]; default 512^3 = 134 M cells), V-cycle only.
    Per kernel class: algorithmic bytes / device time on the fine level, against the measured HBM peak."""
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
    n = args.size
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
    del w, b_host
    build_s = time.perf_counter() - t0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def sync():
        torch.cuda.synchronize()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        solver.applyVCycleDevice(Z, B)
    sampler = ClockSampler(local_rank)
    sync()
    sampler.start()
    ctx.launch_count(reset=True)
    vc = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_begin()
        solver.applyVCycleDevice(Z, B)
        vc.append(ctx.timer_end())
    sync()
    launches, comm_ops = ctx.launch_count(), ctx.comm_count()
    clocks = sampler.stop()
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
    prof_all, prof_fine = ctx.profile(False), ctx.profile(True)
    ctx.profile_enable(False)
    peak, peak_kind = measured_peak()
    fine = {k: v for k, v in prof_fine.items() if k not in ("setup", "coarse_solve", "halo_exchange") and v[1] > 0 and v[2] > 0}
    classes = {k: {"us_per_launch": v[0] / v[1] * 1e3, "launches_per_vcycle": v[1] // nprof, "algorithmic_gbs": v[2] / (v[0] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": v[2] / (v[0] * 1e-3) / 1e9 / peak} for k, v in fine.items()}
    dom = max(fine, key=lambda k: fine[k][0])
    d_ms, d_n, d_bytes = fine[dom]
    achieved = d_bytes / (d_ms * 1e-3) / 1e9
    if rank == 0:
        vb = BYTES_VCYCLE_PER_CELL * active
        line = {
            "metric": "vcycle_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{n}^3 fully liquid box (one DIRICHLET layer), expanded {2 * n}^3, one V-cycle, Jacobi smoother", "levels": solver.getMGLevels(),
                       "active_cells": active, "l2": "256 MB flush write before every timed step",
                       "parallelism": "single GPU" if world == 1 else f"{world} z-slabs over NCCL"},
            "vcycle_algorithmic_gbs": vb / (ms * 1e-3) / 1e9, "vcycle_frac_of_hbm_peak": vb / (ms * 1e-3) / 1e9 / peak / world,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(f"vcycle{n}", dom),
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "launches": d_n, "avg_launch_us": d_ms / d_n * 1e3,
                         "algorithmic_bytes_per_launch": d_bytes / d_n},
            "fine_level_kernels": classes, "kernel_timing": "CUDA event nodes inside the replayed V-cycle graph (rank 0's slab when sharded)",
            "clocks": clocks, "gpu_launches": int(launches), "nccl_ops": int(comm_ops), "host_build_s": build_s,
        }
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
    ap.add_argument("--workload", default="pcg", choices=["pcg", "vcycle"], help="pcg: the headline 256^3 MGPCG solve; vcycle: V-cycle-only sweep (config 4)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.size is None:
        args.size = 512 if args.workload == "vcycle" else 256
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    elif args.workload == "vcycle":
        run_vcycle_sweep(args, rank, world, local_rank)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
