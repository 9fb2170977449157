"""CPU: the z-slab partition behind the sharded solver (gmg_shard_plan, a pure host function of the C ABI) and the
two-rank halo bookkeeping, replayed with torch.distributed/gloo on the CPU oracle's operators."""
import os
import sys

import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import api


def storage_levels(base_n, levels):
    """z planes / fine->coarse shift / cells of the cropped storage boxes for a cubic base grid (csrc makeGeom)."""
    pad = 2 ** (levels - 1)
    lo, hi = pad, pad + base_n
    planes, shift, cells, orgs = [], [], [], []
    for l in range(levels):
        org = 2 * (lo // 2 - 1)
        end = 2 * (-(-hi // 2) + 1)
        n = end - org
        planes.append(n)
        orgs.append(org)
        cells.append(n * n * (-(-n // 16) * 16))
        lo, hi = lo // 2, -(-hi // 2)
    for l in range(levels):
        if l + 1 < levels:
            shift.append(orgs[l] // 2 - orgs[l + 1])
        else:
            shift.append(0)
    return planes, shift, cells


@pytest.mark.parametrize("base_n,levels,world", [(256, 7, 2), (256, 7, 4), (256, 7, 8), (512, 8, 8), (64, 5, 2), (128, 6, 3)])
def test_cuts_nest_and_respect_halo_depth(base_n, levels, world):
    planes, shift, cells = storage_levels(base_n, levels)
    S, cuts = api.shard_plan(planes, shift, cells, world, max_shard_levels=3, min_cells=1000)
    assert 0 <= S <= 3
    for l in range(S):
        c = cuts[l]
        assert c[0] == 0 and c[-1] == planes[l]
        need = 10 if l == 0 else 8
        for k in range(world):
            assert c[k] % 2 == 0 and c[k + 1] - c[k] >= need
        if l + 1 < S:
            for k in range(1, world):
                # both children planes of a coarse plane sit on the same rank as the coarse plane
                assert (c[k] >> 1) + shift[l] == cuts[l + 1][k]
    if base_n >= 256 and world <= 4:
        assert S >= 2


def test_small_boxes_stay_replicated():
    planes, shift, cells = storage_levels(32, 4)
    S, cuts = api.shard_plan(planes, shift, cells, 8, max_shard_levels=3, min_cells=1000)
    assert S == 0  # 36 planes cannot give 8 ranks 10 planes each
    S, _ = api.shard_plan(planes, shift, cells, 2, max_shard_levels=3, min_cells=10**9)
    assert S == 0
    S, _ = api.shard_plan(planes, shift, cells, 1)
    assert S == 0
