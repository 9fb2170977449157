"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun).  Every rank builds the same synthetic domain,
shards the solver into z-slabs and compares each result on its OWNED planes against an unsharded solver of the same
library on the same GPU -- bitwise for everything that involves no cross-rank summation, to 1e-10 for the CG scalars
(the partial sums of the dot products are combined in a different order)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from geometricmultigridpressuresolver_b200 import api  # noqa: E402
from geometricmultigridpressuresolver_b200 import domains as D  # noqa: E402

FAILS = []


def check(cond, msg):
    if not cond:
        FAILS.append(msg)
        print(f"[rank {dist.get_rank()}] FAIL {msg}", flush=True)


def owned(a, lo, hi):
    return a[lo:hi]


def plane_report(a, b):
    d = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
    bad = np.nonzero(d > 0)[0]
    return f"{len(bad)} differing planes of {a.shape[0]}, first {bad[:6].tolist()}, last {bad[-6:].tolist()}, max {d.max():.3e}"


def run_case(ctx_sh, ctx_one, dom, n, shard_levels, backend_nccl=True):
    rank, world = dist.get_rank(), dist.get_world_size()
    os.environ["GMG_SHARD_LEVELS"] = str(shard_levels)
    os.environ["GMG_SHARD_MIN_CELLS"] = "1000"
    bl, bw, dx = D.DOMAINS[dom](n)
    labels, w, off, levels = ctx_one.buildExpandedDomain(bl, bw)
    one = api.GeometricMultigridPoissonSolver(ctx_one, labels, w, levels)
    sh = api.GeometricMultigridPoissonSolver(ctx_sh, labels, w, levels)
    tag = f"{dom}{n} S={shard_levels}"
    info = [sh.shard_info(l) for l in range(sh.getMGLevels())]
    n_sharded = sum(1 for i in info if i[0])
    check(n_sharded == min(shard_levels, n_sharded) and n_sharded >= 1, f"{tag}: expected sharded levels, got {n_sharded}")
    check(sh.getMGLevels() == one.getMGLevels(), f"{tag}: level count")
    if rank == 0:
        print(f"{tag}: levels {sh.getMGLevels()}, sharded {n_sharded}, owned z ranges {[(i[1], i[2]) for i in info[:n_sharded]]}", flush=True)
    # ---- labels (replicated) and band lists (the ranks' owned parts tile the global list)
    for l in range(sh.getMGLevels()):
        check((sh.level_labels(l) == one.level_labels(l)).all(), f"{tag}: labels L{l}")
        mine = sh.level_boundary_cells(l)
        if info[l][0]:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            allc = np.concatenate([p.reshape(-1, 3) for p in parts])
            ref = one.level_boundary_cells(l)
            key = lambda c: np.lexsort((c[:, 0], c[:, 1], c[:, 2]))
            check(len(allc) == len(ref) and (allc[key(allc)] == ref[key(ref)]).all(), f"{tag}: band list L{l} ({len(allc)} vs {len(ref)})")
        else:
            check((mine == one.level_boundary_cells(l)).all(), f"{tag}: band list L{l} (replicated)")
    # ---- single operators, level by level
    for l in range(n_sharded):
        ll = one.level_labels(l)
        lo, hi = info[l][1], info[l][2]
        xa, ba = D.random_active(ll, 11 + l), D.random_active(ll, 23 + l)

        def both(fn):
            outs = []
            for s in (one, sh):
                X, B, R = s.grid(l, xa), s.grid(l, ba), s.grid(l)
                outs.append(fn(s, X, B, R))
            return outs

        def cmp(name, a, b):
            ok = np.array_equal(owned(a, lo, hi), owned(b, lo, hi))
            check(ok, f"{tag}: {name} L{l}: " + (plane_report(owned(a, lo, hi), owned(b, lo, hi)) if not ok else ""))

        a, b = both(lambda s, X, B, R: (s.applyPoissonMatrix(R, X), R.download())[1])
        cmp("applyPoissonMatrix", a, b)
        a, b = both(lambda s, X, B, R: (s.computePoissonResidual(R, X, B), R.download())[1])
        cmp("computePoissonResidual", a, b)
        a, b = both(lambda s, X, B, R: (s.jacobiPoissonSmoother(X, B), X.download())[1])
        cmp("jacobiPoissonSmoother", a, b)
        a, b = both(lambda s, X, B, R: (s.boundaryJacobiPoissonSmoother(X, B, 3), X.download())[1])
        cmp("boundaryJacobiPoissonSmoother x3", a, b)
        d1, d2 = both(lambda s, X, B, R: s.dotProduct(X, B))
        check(abs(d1 - d2) <= 1e-12 * abs(d1), f"{tag}: dotProduct L{l} {d1} vs {d2}")
        d1, d2 = both(lambda s, X, B, R: s.squaredL2Norm(X))
        check(abs(d1 - d2) <= 1e-12 * abs(d1), f"{tag}: squaredL2Norm L{l}")
        d1, d2 = both(lambda s, X, B, R: s.infNorm(X))
        check(d1 == d2, f"{tag}: infNorm L{l}")
        if l + 1 < sh.getMGLevels():
            lc = one.level_labels(l + 1)
            ca = D.random_active(lc, 37 + l)
            outs = []
            for s in (one, sh):
                F, Cg = s.grid(l, xa), s.grid(l + 1)
                s.downsample(Cg, F)
                outs.append(Cg.download())
            clo, chi = (info[l + 1][1], info[l + 1][2]) if info[l + 1][0] else (0, lc.shape[0])
            ok = np.array_equal(outs[0][clo:chi], outs[1][clo:chi])
            check(ok, f"{tag}: downsample L{l}->L{l+1}: " + (plane_report(outs[0][clo:chi], outs[1][clo:chi]) if not ok else ""))
            outs = []
            for s in (one, sh):
                F, Cg = s.grid(l, xa), s.grid(l + 1, ca)
                s.upsampleAndAdd(F, Cg)
                outs.append(F.download())
            cmp("upsampleAndAdd", outs[0], outs[1])
    # ---- V-cycle: bitwise on the owned planes
    l0 = one.level_labels(0)
    lo, hi = info[0][1], info[0][2]
    b = D.random_rhs(labels, dx, 777)
    z1 = one.applyVCycle(np.zeros_like(b), b)
    z2 = sh.applyVCycle(np.zeros_like(b), b)
    ok = np.array_equal(z1[lo:hi], z2[lo:hi])
    check(ok, f"{tag}: applyVCycle: " + (plane_report(z1[lo:hi], z2[lo:hi]) if not ok else ""))
    check(not z2[:lo].any() and not z2[hi:].any(), f"{tag}: applyVCycle wrote outside the owned planes")
    x0 = D.random_active(l0, 5, 1e-3)
    z1 = one.applyVCycle(x0, b, useInitialGuess=True)
    z2 = sh.applyVCycle(x0, b, useInitialGuess=True)
    ok = np.array_equal(z1[lo:hi], z2[lo:hi])
    check(ok, f"{tag}: applyVCycle(useInitialGuess): " + (plane_report(z1[lo:hi], z2[lo:hi]) if not ok else ""))
    # ---- PCG
    x1, it1, h1 = one.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 200)
    x2, it2, h2 = sh.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-6, 200)
    check(it1 == it2, f"{tag}: PCG iterations {it1} vs {it2}")
    if len(h1) == len(h2):
        check((np.abs(h1 - h2) / h1).max() < 1e-9, f"{tag}: PCG history dev {(np.abs(h1 - h2) / h1).max():.3e}")
    else:
        check(False, f"{tag}: PCG history length {len(h1)} vs {len(h2)}")
    check(np.abs(x1[lo:hi] - x2[lo:hi]).max() <= 1e-9 * np.abs(x1).max(), f"{tag}: PCG pressure dev {np.abs(x1[lo:hi] - x2[lo:hi]).max():.3e}")
    full = torch.from_numpy(x2.copy())
    full[:lo] = 0
    full[hi:] = 0
    g = full.cuda()
    dist.all_reduce(g)
    check(np.abs(g.cpu().numpy() - x1).max() <= 1e-9 * np.abs(x1).max(), f"{tag}: PCG pressure assembled over the ranks")
    # plain CG (no preconditioner) exercises the same halo logic without the V-cycle
    # (a short prefix only: unpreconditioned CG amplifies the summation-order difference by ~2.4x per iteration)
    x1, it1, h1 = one.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-12, 12, useMGPreconditioner=False)
    x2, it2, h2 = sh.solveGeometricConjugateGradient(np.zeros_like(b), b, 1e-12, 12, useMGPreconditioner=False)
    check(it1 == it2 and len(h1) == len(h2) and (np.abs(h1 - h2) / h1).max() < 1e-8, f"{tag}: plain CG history {(np.abs(h1 - h2) / h1).max() if len(h1) == len(h2) else -1:.3e}")
    if rank == 0:
        print(f"{tag}: PCG {it2 + 1} iterations, final {h2[-1]:.3e}, comm ops so far {ctx_sh.comm_count()}", flush=True)
    sh.close()
    one.close()


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx_one = api.Context(local)
    ctx_sh = api.Context(local)
    ctx_sh.shard_with_torch(dist)
    # half_box: the lower ranks own no active cell at all (reductions and CG scalars on an empty slab)
    cases = [("sphere", 64, 1), ("sphere", 64, 2), ("flipsplash", 64, 3), ("complex", 64, 2), ("half_box", 64, 2)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:2]
    # every rank must own at least the stored halo depth of planes on a sharded level: larger domains for larger worlds
    scale = 1 if dist.get_world_size() <= 4 else 2
    for dom, n, S in cases:
        run_case(ctx_sh, ctx_one, dom, n * scale, min(S, 2) if scale > 1 else S)
    nf = torch.tensor([len(FAILS)], device="cuda")
    dist.all_reduce(nf)
    if dist.get_rank() == 0:
        print("SHARD_OK" if nf.item() == 0 else f"SHARD_FAIL {int(nf.item())}", flush=True)
    dist.barrier()
    ctx_sh.close()
    ctx_one.close()
    dist.destroy_process_group()
    sys.exit(0 if nf.item() == 0 else 1)


if __name__ == "__main__":
    main()
