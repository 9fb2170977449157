"""CPU: the C-ABI library loads, exports every symbol include/gmg_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from geometricmultigridpressuresolver_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gmg_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmg_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_python_table():
    assert header_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build_product()
    assert os.path.exists(api.LIB_PATH)
    out = subprocess.check_output(["nm", "-D", "--defined-only", api.LIB_PATH], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    lib = api.load_library()
    for s in header_symbols():
        assert getattr(lib, s) is not None
    assert lib.gmg_version() >= 100


def test_no_torch_or_oracle_in_the_product_library():
    out = subprocess.check_output(["ldd", api.LIB_PATH], text=True)
    assert "torch" not in out and "gmg_oracle" not in out and "gmg_ref" not in out


def test_pure_host_entry_points():
    lib = api.load_library()
    base, exp, off, lv = (C.c_int64 * 3)(64, 64, 64), (C.c_int64 * 3)(), (C.c_int64 * 3)(), C.c_int()
    assert lib.gmg_expand_dims(base, exp, off, C.byref(lv)) == 0
    assert list(exp) == [128, 128, 128] and list(off) == [16, 16, 16] and lv.value == 5
    opt = api.SolverOptions()
    lib.gmg_solver_default_options(C.byref(opt))
    assert opt.boundary_width == 3 and opt.boundary_iterations == 3 and opt.coarse_matrix_scale == 1.0 and opt.use_gauss_seidel == 0
    assert lib.gmg_kernel_class_count() >= 8
    assert lib.gmg_kernel_class_name(0) == b"jacobi_interior"


def test_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device, context creation is an error, not a slow path."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(api.GmgError) as e:
        api.Context(0)
    assert e.value.status == 1 and "no CPU fallback" in str(e.value)
