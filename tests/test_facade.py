"""The C++ host facade (include/gmg_b200_hdk.hpp): compiles against the HDK container surface, fails loudly without a GPU
(CPU checks), and matches the reference's own sources call for call on the B200 (GPU check, oracle/_ref/test_facade)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from geometricmultigridpressuresolver_b200 import domains as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "test_facade")


def write_case(path, dom, n):
    labels, w, dx = D.DOMAINS[dom](n)
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", labels.shape[2], labels.shape[1], labels.shape[0]))
        f.write(np.ascontiguousarray(labels, dtype=np.int32).tobytes())
        for a in range(3):
            f.write(np.ascontiguousarray(w[a], dtype=np.float64).tobytes())
        f.write(struct.pack("<d", dx))


def test_facade_header_compiles_against_the_hdk_surface(tmp_path):
    """Syntax/semantic check of the facade in its drop-in namespace (HDK) with every template instantiated."""
    src = tmp_path / "inst.cpp"
    src.write_text(
        """
#include "gmg_b200_hdk.hpp"
namespace Ops = HDK::GeometricMultigridOperators;
using W = std::array<UT_VoxelArray<double>, 3>;
void instantiate(UT_VoxelArray<double> &x, const UT_VoxelArray<double> &b, UT_VoxelArray<int> &l, W &w, UT_Array<UT_Vector3I> &cells)
{
    auto f = [](int v) { return v == 0; };
    Ops::buildExpandedCellLabels(l, l, f, f, f);
    Ops::buildExpandedBoundaryWeights(w[0], w[0], l, UT_Vector3I(0, 0, 0), 0);
    Ops::setBoundaryCellLabels(l, w);
    l = Ops::buildCoarseCellLabels(l);
    cells = Ops::buildBoundaryCells(l, 3);
    Ops::jacobiPoissonSmoother<double>(x, b, l, &w);
    Ops::boundaryJacobiPoissonSmoother<double>(x, b, l, cells, &w);
    Ops::applyPoissonMatrix<double>(x, b, l, &w);
    Ops::computePoissonResidual<double>(x, x, b, l);
    Ops::downsample<double>(x, b, l, l);
    Ops::upsampleAndAdd<double>(x, b, l, l);
    Ops::addToVector<double>(x, b, 1.0, l);
    Ops::addVectors<double>(x, b, b, 1.0, l);
    Ops::scaleVector<double>(x, 2.0, l);
    (void)Ops::dotProduct<double>(x, b, l);
    (void)Ops::squaredL2Norm<double>(x, l);
    (void)Ops::l2Norm<double>(x, l);
    (void)Ops::infNorm(x, l);
    Ops::uncompressActiveGrid(x, l);
    HDK::GeometricMultigridPoissonSolver mg(l, w, 3, false);
    mg.applyVCycle(x, b);
    (void)mg.getMGLevels();
    (void)HDK::solveGeometricConjugateGradient(mg, x, b, 1e-5, 100);
}
"""
    )
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"), str(src)])


def test_facade_binary_fails_loudly_without_a_gpu(tmp_path):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/test_facade is built only where /root/reference exists")
    case = tmp_path / "case.bin"
    write_case(str(case), "sphere", 16)
    p = subprocess.run([BIN, str(case)], capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stdout + p.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("dom,n", [("sphere", 32), ("complex", 24), ("flipsplash", 32)])
def test_facade_matches_reference_sources(tmp_path, dom, n):
    assert os.path.exists(BIN), "oracle/_ref/test_facade must travel with the repo (built by __graft_entry__.build())"
    case = tmp_path / "case.bin"
    write_case(str(case), dom, n)
    p = subprocess.run([BIN, str(case)], capture_output=True, text=True, timeout=600)
    print(p.stdout[-3000:], p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert "facade parity ok" in p.stdout
