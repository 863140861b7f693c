"""TEST INFRASTRUCTURE ONLY: ctypes access to the two checkers under oracle/.

  RefOracle   -- the reference's own step code compiled from /root/reference
                 (oracle/_ref/libref_LX_LY.so, built by oracle/Makefile from oracle/ref_shim.cu)
  PortOracle  -- the plain-C restatement (oracle/liboracle.so from oracle/psim_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from particle_simulator_b200.frame import METADATA_DTYPE, PARTICLE_DTYPE, FrameBuffer

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_LIB = os.path.join(HERE, "liboracle.so")
REFERENCE_ROOT = "/root/reference"


def build(grids: tuple[str, ...] | None = None) -> None:
    """Compile the restatement, and the reference itself where /root/reference exists."""
    cmd = ["make", "-C", HERE, "oracle", "ref"]
    if grids:
        cmd.append("GRIDS=" + " ".join(grids))
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def ref_lib_path(lx: int, ly: int) -> str:
    return os.path.join(REF_DIR, f"libref_{lx}_{ly}.so")


def ref_available(lx: int = 6, ly: int = 6) -> bool:
    return os.path.exists(ref_lib_path(lx, ly))


class CGrid(ctypes.Structure):
    _fields_ = [("lx", ctypes.c_uint32), ("ly", ctypes.c_uint32), ("capacity", ctypes.c_uint32)]


class CMie(ctypes.Structure):
    _fields_ = [("sigma", ctypes.c_float), ("epsilon", ctypes.c_float), ("n", ctypes.c_float), ("m", ctypes.c_float)]


def _mie(meta: np.ndarray) -> CMie:
    p = meta["particles"][0]
    return CMie(float(p["sigma"]), float(p["epsilon"]), float(p["n"]), float(p["m"]))


class RefOracle:
    """The compiled reference for one grid size. Loading allocates 3 slot arrays (kernel.cuh:42-66),
    so load the big grids only when needed. One instance per grid per process (the reference keeps
    its state in a global `Kernel kernel;`, kernel.cuh:134)."""

    _loaded: dict[tuple[int, int], "RefOracle"] = {}

    def __new__(cls, lx: int = 6, ly: int = 6):
        key = (lx, ly)
        if key not in cls._loaded:
            self = super().__new__(cls)
            self._init(lx, ly)
            cls._loaded[key] = self
        return cls._loaded[key]

    def _init(self, lx: int, ly: int) -> None:
        path = ref_lib_path(lx, ly)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build it with `make -C oracle ref GRIDS={lx}_{ly}` "
                                    f"(needs {REFERENCE_ROOT})")
        L = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        for name in ("ref_grid_x_log2", "ref_grid_y_log2", "ref_bucket_capacity", "ref_slot_count",
                     "ref_hardware_threads"):
            getattr(L, name).restype = ctypes.c_uint32
            getattr(L, name).argtypes = []
        L.ref_gpu_count.restype = ctypes.c_int
        L.ref_prepare.restype = ctypes.c_uint32
        L.ref_prepare.argtypes = [vp]
        L.ref_set_metadata.restype = None
        L.ref_set_metadata.argtypes = [vp]
        L.ref_read_slots.restype = None
        L.ref_read_slots.argtypes = [vp]
        L.ref_read_compact.restype = None
        L.ref_read_compact.argtypes = [vp]
        L.ref_move.restype = None
        L.ref_step.restype = None
        L.ref_run_frame.restype = ctypes.c_double
        L.ref_compact_step.restype = None
        L.ref_compact_step.argtypes = [vp, vp, vp, ctypes.c_uint32]
        L.ref_params_C.restype = ctypes.c_float
        L.ref_params_C.argtypes = [CMie]
        L.ref_f_force.restype = ctypes.c_float
        L.ref_f_force.argtypes = [CMie, ctypes.c_float]
        self.L = L
        self.lx, self.ly = int(L.ref_grid_x_log2()), int(L.ref_grid_y_log2())
        assert (self.lx, self.ly) == (lx, ly)
        self.capacity = int(L.ref_bucket_capacity())
        self.slot_count = int(L.ref_slot_count())
        self.gpu_count = int(L.ref_gpu_count())
        self.hardware_threads = int(L.ref_hardware_threads())

    def prepare(self, frame: FrameBuffer) -> int:
        """kernel_prepare_frame + Kernel::write. Returns the device actually used."""
        return int(self.L.ref_prepare(frame.ptr))

    def set_metadata(self, meta: np.ndarray) -> None:
        m = np.ascontiguousarray(meta, dtype=METADATA_DTYPE)
        self.L.ref_set_metadata(ctypes.c_void_p(m.ctypes.data))

    def slots(self) -> np.ndarray:
        out = np.zeros(self.slot_count, dtype=PARTICLE_DTYPE)
        self.L.ref_read_slots(ctypes.c_void_p(out.ctypes.data))
        return out

    def compact(self) -> FrameBuffer:
        fb = FrameBuffer(self.slot_count)
        fb.count = fb.capacity
        self.L.ref_read_compact(fb.ptr)
        return fb

    def move(self) -> None:
        self.L.ref_move()

    def step(self) -> None:
        self.L.ref_step()

    def run_frame(self) -> float:
        return float(self.L.ref_run_frame())

    def compact_step(self, particles: np.ndarray, meta: np.ndarray) -> np.ndarray:
        src = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        dst = np.zeros_like(src)
        m = np.ascontiguousarray(meta, dtype=METADATA_DTYPE)
        self.L.ref_compact_step(ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data),
                                ctypes.c_void_p(m.ctypes.data), len(src))
        return dst

    def params_C(self, meta: np.ndarray) -> float:
        return float(self.L.ref_params_C(_mie(meta)))

    def f_force(self, meta: np.ndarray, r: float) -> float:
        return float(self.L.ref_f_force(_mie(meta), r))


class PortOracle:
    """The plain-C restatement, any grid size."""

    _lib = None

    def __init__(self, lx: int = 6, ly: int = 6, capacity: int = 16):
        if PortOracle._lib is None:
            if not os.path.exists(PORT_LIB):
                build()
            L = ctypes.CDLL(PORT_LIB)
            vp = ctypes.c_void_p
            L.oracle_slot_count.restype = ctypes.c_uint64
            L.oracle_slot_count.argtypes = [CGrid]
            L.oracle_params_C.restype = ctypes.c_float
            L.oracle_params_C.argtypes = [CMie]
            L.oracle_f_force.restype = ctypes.c_float
            L.oracle_f_force.argtypes = [CMie, ctypes.c_float]
            L.oracle_prepare.restype = ctypes.c_uint32
            L.oracle_prepare.argtypes = [vp, vp, CGrid]
            L.oracle_move.restype = ctypes.c_uint64
            L.oracle_move.argtypes = [vp, vp, CGrid]
            L.oracle_step.restype = None
            L.oracle_step.argtypes = [vp, vp, vp, CGrid, ctypes.c_uint32]
            L.oracle_step_species.restype = None
            L.oracle_step_species.argtypes = [vp, vp, vp, CGrid, ctypes.c_uint32]
            L.oracle_forces.restype = None
            L.oracle_forces.argtypes = [vp, vp, CGrid, vp, vp, vp]
            L.oracle_run_frame.restype = ctypes.c_uint32
            L.oracle_run_frame.argtypes = [vp, vp, vp, vp, CGrid, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
            L.oracle_compact.restype = None
            L.oracle_compact.argtypes = [vp, vp, CGrid, vp]
            L.oracle_diagnostics.restype = None
            L.oracle_diagnostics.argtypes = [vp, vp, CGrid, vp]
            PortOracle._lib = L
        self.L = PortOracle._lib
        self.grid = CGrid(lx, ly, capacity)
        self.lx, self.ly, self.capacity = lx, ly, capacity
        self.slot_count = int(self.L.oracle_slot_count(self.grid))

    @staticmethod
    def _p(a: np.ndarray) -> ctypes.c_void_p:
        return ctypes.c_void_p(a.ctypes.data)

    def _meta(self, meta: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(meta, dtype=METADATA_DTYPE)

    def params_C(self, meta: np.ndarray) -> float:
        return float(self.L.oracle_params_C(_mie(meta)))

    def f_force(self, meta: np.ndarray, r: float) -> float:
        return float(self.L.oracle_f_force(_mie(meta), r))

    def prepare(self, frame: FrameBuffer) -> tuple[np.ndarray, int]:
        slots = np.zeros(self.slot_count, dtype=PARTICLE_DTYPE)
        dropped = int(self.L.oracle_prepare(frame.ptr, self._p(slots), self.grid))
        return slots, dropped

    def move(self, slots: np.ndarray) -> tuple[np.ndarray, int]:
        dst = np.zeros_like(slots)
        live = int(self.L.oracle_move(self._p(slots), self._p(dst), self.grid))
        return dst, live

    def step(self, slots: np.ndarray, meta: np.ndarray, threads: int = 1) -> np.ndarray:
        dst = np.zeros_like(slots)
        m = self._meta(meta)
        self.L.oracle_step(self._p(slots), self._p(dst), self._p(m), self.grid, threads)
        return dst

    def step_species(self, slots: np.ndarray, meta: np.ndarray, threads: int = 1) -> np.ndarray:
        """The per-species extension (oracle_step_species): not reference behaviour, the spec of PsimConfig.species_physics."""
        dst = np.zeros_like(slots)
        m = self._meta(meta)
        self.L.oracle_step_species(self._p(slots), self._p(dst), self._p(m), self.grid, threads)
        return dst

    def forces(self, slots: np.ndarray, meta: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        fx = np.zeros(self.slot_count, dtype=np.float32)
        fy = np.zeros_like(fx)
        mp = np.zeros_like(fx)
        m = self._meta(meta)
        self.L.oracle_forces(self._p(slots), self._p(m), self.grid, self._p(fx), self._p(fy), self._p(mp))
        return fx, fy, mp

    def run_frame(self, slots: np.ndarray, meta: np.ndarray, threads: int = 1) -> tuple[np.ndarray, int, int]:
        """Returns (result slots, steps executed, moves)."""
        b0 = np.ascontiguousarray(slots).copy()
        b1 = np.zeros_like(b0)
        b2 = np.zeros_like(b0)
        m = self._meta(meta)
        moves = ctypes.c_uint32()
        steps = int(self.L.oracle_run_frame(self._p(b0), self._p(b1), self._p(b2), self._p(m), self.grid, threads,
                                            ctypes.byref(moves)))
        return b1, steps, int(moves.value)

    def compact(self, slots: np.ndarray, meta: np.ndarray) -> FrameBuffer:
        live = int((slots["ty"] >= 0).sum())
        fb = FrameBuffer(max(live, 1))
        m = self._meta(meta)
        self.L.oracle_compact(self._p(slots), self._p(m), self.grid, fb.ptr)
        return fb

    def diagnostics(self, slots: np.ndarray, meta: np.ndarray) -> dict[str, float]:
        out = np.zeros(6, dtype=np.float64)
        m = self._meta(meta)
        self.L.oracle_diagnostics(self._p(slots), self._p(m), self.grid, self._p(out))
        return dict(ke=out[0], pe_pair=out[1], pe_wall=out[2], px=out[3], py=out[4], live=int(out[5]))
