"""CPU, world_size 2 over gloo: the deep-halo schedule of the z-slab sharding (csrc/gmg_b200.cu vcycleLaunches / pcgDevice,
DESIGN.md section 6) replayed with the CPU oracle's operators.

Each rank keeps full-size arrays but only trusts its slab: every active cell outside the planes it is entitled to know is
poisoned with NaN, so a stencil that reaches one plane too far -- an exchange that is too shallow, a sweep too many between
exchanges -- shows up as NaN (or as a value that differs from the unsharded oracle) on the rank's OWNED planes.  The cut
planes come from the library's own partition function (gmg_shard_plan, host-only)."""
import os
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HALO_X, HALO_P = 8, 9  # csrc/gmg_common.cuh


def _geometry(labels_per_level):
    """z extent / origin / fine->coarse shift of the library's cropped storage boxes (csrc makeGeom, z axis only)."""
    l0 = labels_per_level[0]
    zs = np.nonzero((l0 != 1).any(axis=(1, 2)))[0]
    lo, hi = int(zs[0]), int(zs[-1]) + 1
    org, planes = [], []
    for _ in labels_per_level:
        o = 2 * (lo // 2 - 1)
        e = 2 * (-(-hi // 2) + 1)
        org.append(o)
        planes.append(e - o)
        lo, hi = lo // 2, -(-hi // 2)
    shift = [org[l] // 2 - org[l + 1] for l in range(len(org) - 1)] + [0]
    return org, planes, shift


def _worker(rank, world, port_file, result_q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist

        from geometricmultigridpressuresolver_b200 import api
        from geometricmultigridpressuresolver_b200 import domains as D
        from oracle import bindings

        dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
        port = bindings.PortLib()
        bl, bw, dx = D.complex_domain(32)  # 4 levels, free-surface weights at level 0
        labels, w, off, levels = port.expand_domain(bl, bw)
        full = port.solver(labels, w, levels, False)
        nl = full.levels
        lab = [full.level_labels(l) for l in range(nl)]
        cells = [full.level_boundary_cells(l) for l in range(nl)]
        act = [(a == 0) | (a == 3) for a in lab]
        org, planes, shift = _geometry(lab)
        S, cuts = api.shard_plan(planes, shift, [10**6] * nl, world, max_shard_levels=2, min_cells=1000)
        assert S == 2, (S, planes)
        own = [(0 if rank == 0 else org[l] + cuts[l][rank], lab[l].shape[0] if rank == world - 1 else org[l] + cuts[l][rank + 1]) for l in range(S)]
        gather = [(0 if k == 0 else (org[S - 1] + cuts[S - 1][k]) // 2, lab[S].shape[0] if k == world - 1 else (org[S - 1] + cuts[S - 1][k + 1]) // 2)
                  for k in range(world)]
        ones = [np.ones(api.face_shape(lab[S].shape, a)) for a in range(3)]
        sub = port.solver(lab[S], ones, nl - S, False)  # the replicated levels, unsharded

        def keep(a, l, zlo, zhi):
            out = a.copy()
            m = act[l].copy()
            m[max(zlo, 0):max(zhi, 0)] = False
            out[m] = np.nan
            return out

        def exchange(a, l, depth):
            """refresh `depth` planes beyond the owned slab from the z-neighbour (2 ranks)"""
            lo, hi = own[l]
            peer = 1 - rank
            mine = a[hi - depth:hi] if rank == 0 else a[lo:lo + depth]
            send = torch.from_numpy(np.ascontiguousarray(mine))
            recv = torch.empty_like(send)
            if rank == 0:
                dist.send(send, peer); dist.recv(recv, peer)
                a[hi:hi + depth] = recv.numpy()
            else:
                dist.recv(recv, peer); dist.send(send, peer)
                a[lo - depth:lo] = recv.numpy()
            return a

        def smooth(l, x, b, wts):
            x = port.boundary_jacobi(x, b, lab[l], cells[l], 3, wts)
            x = port.jacobi(x, b, lab[l], wts)
            return port.boundary_jacobi(x, b, lab[l], cells[l], 3, wts)

        def vcycle(b0):
            """b0: valid HALO_X planes deep.  Returns x valid on the owned planes (1 deep, in fact)."""
            b = [b0] + [None] * S
            x = [None] * (S + 1)
            for l in range(S):
                wts = w if l == 0 else None
                x[l] = smooth(l, np.zeros_like(b[l]), b[l], wts)
                r = port.residual(x[l], b[l], lab[l], wts)
                bc = port.downsample(r, lab[l + 1], lab[l])
                if l + 1 < S:
                    bc = keep(bc, l + 1, *own[l + 1])
                    bc = exchange(bc, l + 1, HALO_X)
                else:
                    bc = keep(bc, l + 1, *gather[rank])
                    parts = [torch.zeros(1)] * world
                    t = torch.from_numpy(np.ascontiguousarray(bc[gather[rank][0]:gather[rank][1]]))
                    got = [None] * world
                    dist.all_gather_object(got, t.numpy())
                    for k in range(world):
                        bc[gather[k][0]:gather[k][1]] = got[k]
                b[l + 1] = bc
            if np.isnan(b[S]).any():
                return np.full_like(b0, np.nan)  # a too shallow halo already poisoned the replicated levels
            x[S] = sub.vcycle(np.zeros_like(b[S]), b[S])
            for l in range(S - 1, -1, -1):
                wts = w if l == 0 else None
                x[l] = port.upsample_add(x[l], x[l + 1], lab[l], lab[l + 1])
                x[l] = keep(x[l], l, *own[l])
                x[l] = exchange(x[l], l, HALO_X)
                x[l] = smooth(l, x[l], b[l], wts)
            return x[0]

        lo, hi = own[0]
        b = D.random_rhs(labels, dx, 99)
        # ---- V-cycle: bitwise equal to the unsharded oracle on the owned planes
        z_ref = full.vcycle(np.zeros_like(b), b)
        z = vcycle(keep(b, 0, lo - HALO_X, hi + HALO_X))
        assert not np.isnan(z[lo:hi]).any(), "NaN reached the owned planes: halo too shallow"
        assert np.array_equal(z[lo:hi], z_ref[lo:hi]), "sharded V-cycle differs from the unsharded one"
        # the schedule is tight: one plane less of right-hand side must break it
        z_bad = vcycle(keep(b, 0, lo - (HALO_X - 1), hi + (HALO_X - 1)))
        assert np.isnan(z_bad[lo:hi]).any(), "a 7-plane rhs halo should not be enough"
        # ---- three PCG iterations with the p-exchange 9 deep; r is never exchanged
        x_ref, it_ref, hist_ref = full.pcg(np.zeros_like(b), b, 1e-30, 3)

        def allsum(v):
            t = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        def dot_owned(a, c):
            return allsum(float(np.sum((a[lo:hi] * c[lo:hi])[act[0][lo:hi]])))

        r = keep(b, 0, lo - HALO_X, hi + HALO_X)  # x0 = 0: r = b, valid 8 deep
        bb = dot_owned(r, r)
        x = np.zeros_like(b)
        p = vcycle(r)
        rho = dot_owned(p, r)
        hist = []
        for _ in range(3):
            p = keep(p, 0, lo, hi)
            p = exchange(p, 0, HALO_P)
            t = port.apply(p, lab[0], w)          # valid 8 deep
            alpha = rho / dot_owned(p, t)
            x = x + alpha * p
            r = r - alpha * t                      # stays valid 8 deep
            assert not np.isnan(r[max(lo - HALO_X, 0):hi + HALO_X][act[0][max(lo - HALO_X, 0):hi + HALO_X]]).any(), "r lost its halo"
            hist.append(np.sqrt(dot_owned(r, r) / bb))
            zz = vcycle(r)
            rho_new = dot_owned(zz, r)
            p = zz + (rho_new / rho) * p
            rho = rho_new
        assert np.allclose(hist, hist_ref[:3], rtol=1e-10, atol=0), (hist, hist_ref[:3])
        result_q.put((rank, "ok"))
    except Exception:
        result_q.put((rank, traceback.format_exc()))


def test_two_rank_deep_halo_schedule(tmp_path):
    import torch.multiprocessing as mp

    from oracle import bindings

    bindings.build(ref=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_file = str(tmp_path / "rendezvous")
    procs = [ctx.Process(target=_worker, args=(r, 2, port_file, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
