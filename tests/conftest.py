import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build the CPU-side libraries (and the oracle restatement) once per session."""
    from particle_simulator_b200 import _build

    _build.build_io()
    import subprocess

    subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "oracle"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def frame_from(particles: np.ndarray, meta: np.ndarray, capacity: int | None = None):
    from particle_simulator_b200 import FrameBuffer

    fb = FrameBuffer(max(len(particles), 1) if capacity is None else capacity, np.asarray(meta).reshape(()))
    fb.set_particles(particles)
    return fb


@pytest.fixture(scope="session")
def golden():
    cache: dict[str, dict] = {}

    def get(name: str) -> dict:
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get
