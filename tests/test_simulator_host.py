"""Host-side behaviour of the simulator process that needs no GPU: its command line, and that it fails loudly
(no CPU path) when there is no B200 -- before it touches the network."""
import subprocess

import pytest

from particle_simulator_b200 import _build


@pytest.fixture(scope="module")
def binary():
    _build.build_all()
    return _build.SIMULATOR


def test_help_and_bad_arguments(binary):
    p = subprocess.run([binary, "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and "drop-in for cuda_simulator" in p.stderr
    p = subprocess.run([binary, "--no-such-flag"], capture_output=True, text=True)
    assert p.returncode == 2 and "usage:" in p.stderr
    p = subprocess.run([binary, "--grid", "6"], capture_output=True, text=True)
    assert p.returncode == 2


def test_without_a_gpu_it_refuses_to_start(binary):
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    p = subprocess.run([binary, "--connect", "127.0.0.1:9"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 3
    assert "no CPU path" in p.stderr
