"""Slab decomposition with one process per GPU and NCCL as the transport (needs >= 2 GPUs; on a one-GPU
box the same slabs are covered by tests/test_gpu_slabs.py through the in-process group)."""
import os
import signal
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("halo", ["push", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_slabs_are_bit_identical_to_single_slab(world, halo):
    """halo = push: the step kernel stores boundary rows into the neighbours' ghost rows over peer memory (CUDA IPC);
    halo = nccl: an ncclSend/ncclRecv pair per neighbour after every step (PSIM_HALO=nccl)."""
    run_workers(world, halo)


def test_nccl_slabs_of_unequal_heights():
    """PsimConfig.slab_bounds across processes: slab 0 owns 448 of the 1024 cell rows (38 % of the particles), slab 1 the rest."""
    run_workers(2, "push", bounds="0,448,1024", port=29650)


def test_nccl_slabs_metadata_change_between_frames():
    """New sigma / box between frames while the halo is pushed: own rows rebuilt locally, ghost rows re-delivered."""
    run_workers(2, "push", port=29660, meta_change=True)


@pytest.mark.parametrize("world", [2, 4])
def test_rebalance_across_processes(world):
    """slabs.rebalance_across_ranks: boundaries moved between frames of a running one-process-per-GPU decomposition."""
    run_workers(world, "push", port=29670 + world, script="mp_rebalance_worker.py")


def run_workers(world, halo, bounds=None, port=None, meta_change=False, script="mp_slab_worker.py"):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpus()}")
    env = dict(os.environ, PSIM_EXPECT_HALO=halo)
    if bounds:
        env["PSIM_TEST_BOUNDS"] = bounds
    if halo == "nccl":
        env["PSIM_HALO"] = "nccl"
    if meta_change:
        env["PSIM_TEST_META_CHANGE"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port or 29600 + world + (10 if halo == "nccl" else 0)),
           os.path.join(REPO, "tests", script), "3"]
    # own session: if a rank hangs, the whole process group is killed, never left spinning on the GPUs
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=REPO, env=env,
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("NCCL slab worker timed out\n" + out[-3000:] + err[-3000:])
    sys.stdout.write(out[-4000:])
    assert proc.returncode == 0, out[-4000:] + err[-4000:]
    assert "identical to the single-slab run: True" in out and "run: False" not in out
