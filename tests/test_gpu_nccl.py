"""Slab decomposition with one process per GPU and NCCL as the transport (needs >= 2 GPUs; on a one-GPU
box the same slabs are covered by tests/test_gpu_slabs.py through the in-process group)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_slabs_are_bit_identical_to_single_slab(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpus()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(REPO, "tests", "mp_slab_worker.py"), "3"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    sys.stdout.write(proc.stdout[-4000:])
    assert proc.returncode == 0, proc.stdout[-4000:] + proc.stderr[-4000:]
    assert "identical to the single-slab run: True" in proc.stdout
