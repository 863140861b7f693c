"""Build recipes for the native libraries (in-tree, sm_100a only).

  libparticle_io_c.so  csrc/particle_io.cpp + csrc/scene.cpp   (g++)   include/particle_io.h, psim_scene.h
  libpsim_b200.so      csrc/stepper.cu                         (nvcc)  include/psim_b200.h

`nvcc` cross-compiles for sm_100a without a GPU. The shared objects are git-ignored but travel to
the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO, "include")
LIB_IO = os.path.join(PKG_DIR, "libparticle_io_c.so")
LIB_PSIM = os.path.join(PKG_DIR, "libpsim_b200.so")
SIMULATOR = os.path.join(PKG_DIR, "psim_simulator")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build_io(force: bool = False) -> str:
    src = [os.path.join(CSRC, "particle_io.cpp"), os.path.join(CSRC, "scene.cpp")]
    deps = src + [os.path.join(INCLUDE, "particle_io.h"), os.path.join(INCLUDE, "psim_scene.h")]
    if force or _stale(LIB_IO, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I" + INCLUDE, *src, "-o", LIB_IO, "-lpthread"])
    return LIB_IO


def build_psim(force: bool = False, verbose: bool = False) -> str:
    src = [os.path.join(CSRC, "stepper.cu")]
    kernels = [os.path.join(CSRC, f) for f in ("device_common.cuh", "step_int.cuh", "step_float.cuh", "binning.cuh")]
    deps = src + kernels + [os.path.join(INCLUDE, "psim_b200.h"), os.path.join(INCLUDE, "particle_io.h")]
    if force or _stale(LIB_PSIM, deps):
        cmd = [_nvcc(), *NVCC_ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-I" + INCLUDE, *src, "-o", LIB_PSIM, "-ldl"]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        _run(cmd)
    return LIB_PSIM


def build_simulator(force: bool = False) -> str:
    """The simulator process (drop-in for the reference's cuda_simulator binary): plain C++ over the two C ABIs."""
    src = [os.path.join(CSRC, "simulator_main.cpp")]
    deps = src + [os.path.join(INCLUDE, "psim_b200.h"), os.path.join(INCLUDE, "particle_io.h"), LIB_IO, LIB_PSIM]
    if force or _stale(SIMULATOR, deps):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-I" + INCLUDE, *src, "-o", SIMULATOR, "-L" + PKG_DIR,
              "-lpsim_b200", "-lparticle_io_c", "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return SIMULATOR


def build_all(force: bool = False) -> None:
    build_io(force)
    build_psim(force)
    build_simulator(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB_IO)
    print(LIB_PSIM)
    print(SIMULATOR)
