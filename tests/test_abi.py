"""Every function declared in include/*.h is exported by the library that header belongs to, and the
product library really is CUDA-only. No compute calls here (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

from particle_simulator_b200 import _build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = {
    "particle_io.h": "io",
    "psim_scene.h": "io",
    "psim_b200.h": "psim",
}


def declared_functions(header: str) -> list[str]:
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"#[^\n]*", "", text)
    text = re.sub(r"typedef\s+(struct|enum)\s+\w+\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", text)))


@pytest.fixture(scope="module")
def libs():
    return {"io": ctypes.CDLL(_build.build_io()), "psim": ctypes.CDLL(_build.build_psim())}


@pytest.mark.parametrize("header", sorted(HEADERS))
def test_header_symbols_are_exported(header, libs):
    names = declared_functions(header)
    assert len(names) >= 4, names
    lib = libs[HEADERS[header]]
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_particle_io_has_the_15_reference_functions():
    want = {"new_tcp_client", "reader_open_file", "reader_destroy", "reader_read", "reader_read_last",
            "writer_open_file", "writer_destroy", "writer_write", "frame_destroy", "frame_print", "frame_compact",
            "frame_compact_into", "packet_size", "frame_header_init", "particle_is_null"}
    assert set(declared_functions("particle_io.h")) == want


def test_headers_compile_as_c_and_cpp(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "particle_io.h"\n#include "psim_scene.h"\n#include "psim_b200.h"\n'
                   "int main(void) { return (int)sizeof(FrameHeader) - 96 + (int)sizeof(PsimConfig) - 72; }\n")
    inc = os.path.join(REPO, "include")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(tmp_path / "tc")], check=True)
    assert subprocess.run([str(tmp_path / "tc")]).returncode == 0
    cpp = tmp_path / "t.cpp"
    cpp.write_text(src.read_text())
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", inc, str(cpp), "-o", str(tmp_path / "tcpp")], check=True)
    assert subprocess.run([str(tmp_path / "tcpp")]).returncode == 0


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(REPO, "particle_simulator_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert '#include "psim_oracle' not in text and "liboracle" not in text, f


def test_stepper_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _build.build_psim()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\w+)", out))
    assert archs == {"100a"}, out


def test_stepper_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from particle_simulator_b200.stepper import PsimError, Stepper

    with pytest.raises(PsimError, match="no CPU path"):
        Stepper()
