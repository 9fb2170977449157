"""GPU (-m gpu), needs >= 2 devices: z-slab sharding against the unsharded path of the same library.
One process per GPU under torchrun (tests/mgpu_worker.py); skipped on a box with fewer GPUs.  The logs of the round's own runs at
2, 4 and 8 GPUs are profiles/r02_mgpu_worker_n{2,4,8}.log, and `bench.py --gpus N` asserts the same on every run (parity_vs_n1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_unsharded(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(p.stdout[-6000:], p.stderr[-3000:])
    assert p.returncode == 0 and "SHARD_OK" in p.stdout
