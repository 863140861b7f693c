"""ctypes binding of libparticle_io_c.so: the reference's particle_io C API (include/particle_io.h)
and the seeded scene generators (include/psim_scene.h).

Mirrors how the reference's C++ side uses it (cuda_simulator/src/lib/frontend.hpp:10-57).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _build
from .frame import HEADER_DTYPE, FrameBuffer, packet_size

_lib = None


class CFrame(ctypes.Structure):  # particle_io/c_api/src/particle.rs:4-10
    _fields_ = [("ptr", ctypes.c_void_p), ("cap", ctypes.c_size_t), ("len", ctypes.c_size_t)]


class CHandle(ctypes.Structure):  # Reader / Writer: opaque 2 x u64
    _fields_ = [("_raw", ctypes.c_uint64 * 2)]


class CMie(ctypes.Structure):
    _fields_ = [("sigma", ctypes.c_float), ("epsilon", ctypes.c_float), ("n", ctypes.c_float), ("m", ctypes.c_float)]


class CParticle(ctypes.Structure):
    _fields_ = [("x", ctypes.c_uint32), ("y", ctypes.c_uint32), ("vx", ctypes.c_float), ("vy", ctypes.c_float),
                ("ty", ctypes.c_int32)]


class CFrameHeader(ctypes.Structure):
    _fields_ = [("bytes", ctypes.c_uint8 * 96)]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = _build.LIB_IO
        if not os.path.exists(path):
            _build.build_io()
        L = ctypes.CDLL(path)
        vp, cp = ctypes.c_void_p, ctypes.c_char_p
        L.packet_size.restype = ctypes.c_size_t
        L.packet_size.argtypes = [ctypes.c_uint32]
        L.frame_header_init.restype = CFrameHeader
        L.frame_header_init.argtypes = []
        L.particle_is_null.restype = ctypes.c_bool
        L.particle_is_null.argtypes = [CParticle]
        L.frame_destroy.restype = None
        L.frame_destroy.argtypes = [ctypes.POINTER(CFrame)]
        for name in ("frame_print", "frame_compact"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp]
        L.frame_compact_into.restype = None
        L.frame_compact_into.argtypes = [vp, vp]
        L.reader_open_file.restype = None
        L.reader_open_file.argtypes = [ctypes.POINTER(CHandle), cp]
        L.reader_destroy.restype = None
        L.reader_destroy.argtypes = [ctypes.POINTER(CHandle)]
        L.reader_read.restype = CFrame
        L.reader_read.argtypes = [ctypes.POINTER(CHandle)]
        L.reader_read_last.restype = ctypes.c_bool
        L.reader_read_last.argtypes = [ctypes.POINTER(CHandle), ctypes.POINTER(CFrame)]
        L.writer_open_file.restype = None
        L.writer_open_file.argtypes = [ctypes.POINTER(CHandle), cp]
        L.writer_destroy.restype = None
        L.writer_destroy.argtypes = [ctypes.POINTER(CHandle)]
        L.writer_write.restype = ctypes.c_bool
        L.writer_write.argtypes = [ctypes.POINTER(CHandle), vp]
        L.new_tcp_client.restype = ctypes.c_bool
        L.new_tcp_client.argtypes = [ctypes.POINTER(CHandle), ctypes.POINTER(CHandle), cp]
        lattice = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double, ctypes.c_double,
                   ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_uint64]
        L.psim_scene_hex_square.restype = ctypes.c_int
        L.psim_scene_hex_square.argtypes = lattice
        L.psim_scene_square.restype = ctypes.c_int
        L.psim_scene_square.argtypes = lattice
        L.psim_scene_hex_rows.restype = ctypes.c_int
        L.psim_scene_hex_rows.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_uint32, ctypes.c_double, ctypes.c_double, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_uint64]
        L.psim_scene_gas.restype = ctypes.c_int
        L.psim_scene_gas.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_uint64]
        L.psim_force0_r.restype = ctypes.c_double
        L.psim_force0_r.argtypes = [CMie]
        _lib = L
    return _lib


def frame_header_init() -> np.ndarray:
    """`frame_header_init()` as a HEADER_DTYPE scalar array."""
    h = lib().frame_header_init()
    return np.frombuffer(bytes(h.bytes), dtype=HEADER_DTYPE)[0].copy()


def _take(cframe: CFrame) -> FrameBuffer | None:
    """Copy a library-owned frame into a FrameBuffer and give the original back (frame_destroy)."""
    if not cframe.ptr:
        return None
    data = ctypes.string_at(cframe.ptr, cframe.len)
    lib().frame_destroy(ctypes.byref(cframe))
    return FrameBuffer.from_bytes(data)


class Reader:
    """particle_io::Reader through its C API (c_api/src/reader.rs)."""

    def __init__(self, handle: CHandle):
        self._h = handle
        self._open = True

    @classmethod
    def open_file(cls, path: str) -> "Reader":
        h = CHandle()
        lib().reader_open_file(ctypes.byref(h), os.fsencode(path))
        return cls(h)

    def read(self) -> FrameBuffer | None:
        return _take(lib().reader_read(ctypes.byref(self._h)))

    def read_last(self) -> tuple[bool, FrameBuffer | None]:
        f = CFrame()
        ok = lib().reader_read_last(ctypes.byref(self._h), ctypes.byref(f))
        return bool(ok), _take(f)

    def close(self) -> None:
        if self._open:
            lib().reader_destroy(ctypes.byref(self._h))
            self._open = False

    def __del__(self):
        self.close()


class Writer:
    """particle_io::Writer through its C API (c_api/src/writer.rs)."""

    def __init__(self, handle: CHandle):
        self._h = handle
        self._open = True

    @classmethod
    def open_file(cls, path: str) -> "Writer":
        h = CHandle()
        lib().writer_open_file(ctypes.byref(h), os.fsencode(path))
        return cls(h)

    def write(self, frame: FrameBuffer) -> bool:
        return bool(lib().writer_write(ctypes.byref(self._h), frame.ptr))

    def close(self) -> None:
        if self._open:
            lib().writer_destroy(ctypes.byref(self._h))
            self._open = False

    def __del__(self):
        self.close()


def new_tcp_client(addr: str) -> tuple[Reader, Writer] | None:
    r, w = CHandle(), CHandle()
    if not lib().new_tcp_client(ctypes.byref(r), ctypes.byref(w), addr.encode()):
        return None
    return Reader(r), Writer(w)


def frame_compact(frame: FrameBuffer) -> None:
    lib().frame_compact(frame.ptr)


def frame_compact_into(src: FrameBuffer, dst: FrameBuffer) -> None:
    dst.count = dst.capacity  # capacity on entry, live count on return (kernel.cuh:208-209)
    lib().frame_compact_into(src.ptr, dst.ptr)


# -- scenes --------------------------------------------------------------------------------------

def _mie(meta: np.ndarray, species: int) -> CMie:
    p = meta["particles"][species]
    return CMie(float(p["sigma"]), float(p["epsilon"]), float(p["n"]), float(p["m"]))


def force0_r(meta: np.ndarray, species: int = 0) -> float:
    return float(lib().psim_force0_r(_mie(meta, species)))


def scene_hex_square(frame: FrameBuffer, nx: int, ny: int, center: tuple[float, float], distance_factor: float = 1.0,
                     v_min: float = 0.0, v_max: float = 0.0, ty: int = 0, seed: int = 0) -> None:
    rc = lib().psim_scene_hex_square(frame.ptr, frame.capacity, nx, ny, center[0], center[1], distance_factor,
                                     v_min, v_max, ty, seed)
    if rc != 0:
        raise ValueError("hex_square: frame too small or bad arguments")


def scene_hex_rows(frame: FrameBuffer, nx: int, ny: int, rows: tuple[int, int], center: tuple[float, float],
                   distance_factor: float = 1.0, v_min: float = 0.0, v_max: float = 0.0, ty: int = 0,
                   seed: int = 0) -> None:
    """Lattice rows [rows[0], rows[1]) of the nx x ny hex lattice centred on `center` (psim_scene_hex_rows)."""
    rc = lib().psim_scene_hex_rows(frame.ptr, frame.capacity, nx, ny, rows[0], rows[1], center[0], center[1],
                                   distance_factor, v_min, v_max, ty, seed)
    if rc != 0:
        raise ValueError("hex_rows: frame too small or bad arguments")


def scene_square(frame: FrameBuffer, nx: int, ny: int, center: tuple[float, float], distance_factor: float = 1.0,
                 v_min: float = 0.0, v_max: float = 0.0, ty: int = 0, seed: int = 0) -> None:
    rc = lib().psim_scene_square(frame.ptr, frame.capacity, nx, ny, center[0], center[1], distance_factor,
                                 v_min, v_max, ty, seed)
    if rc != 0:
        raise ValueError("square: frame too small or bad arguments")


def scene_gas(frame: FrameBuffer, count: int, margin: float, min_dist: float, v_min: float = 0.0,
              v_max: float = 0.0, ty: int = 0, seed: int = 0) -> None:
    rc = lib().psim_scene_gas(frame.ptr, frame.capacity, count, margin, min_dist, v_min, v_max, ty, seed)
    if rc != 0:
        raise ValueError("gas: frame too small, box too crowded or bad arguments")
