import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (oracle/gmg_oracle.c), built on demand."""
    from oracle import bindings

    bindings.build(ref=False)
    return bindings.PortLib()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled against the shim; only where oracle/_ref was built."""
    from oracle import bindings

    if os.path.isdir("/root/reference/Source"):
        bindings.build(ref=True)
    if not os.path.exists(bindings.REF_SO):
        pytest.skip("oracle/_ref/libgmg_ref.so not present (built only where /root/reference exists)")
    return bindings.RefLib()


@pytest.fixture(scope="session")
def testnode():
    """The reference's own diagnostic node (HDK_TestGeometricMultigrid.cpp over the shim); only where oracle/_ref was built."""
    from oracle import bindings

    if os.path.isdir("/root/reference/Source"):
        bindings.build(ref=True)
    if not os.path.exists(bindings.REF_TESTNODE_SO):
        pytest.skip("oracle/_ref/libgmg_ref_testnode.so not present (built only where /root/reference exists)")
    return bindings.TestNodeLib()


@pytest.fixture(scope="session")
def gpu_ctx():
    from geometricmultigridpressuresolver_b200 import api

    ctx = api.Context(0)  # raises without the CUDA library / device: no fallback
    yield ctx
    ctx.close()
