"""Scene presets: the in-memory scene library of the reference's particle_io crate (particle_io/src/presets.rs:84-154:
`Preset::{to_frame, from_frame}`, `Presets`), behind the C ABI of include/psim_scene.h (libparticle_io_c.so,
csrc/scene.cpp). This module is the ctypes mirror of that interface with the reference's method names. A preset keeps
what a scene is made of -- name, box size, the two species' Mie parameters, the particle list -- and nothing of how it
is stepped (dt, steps per frame, cursor come back as the defaults of a new frame, as in the reference)."""
from __future__ import annotations

import ctypes

from . import io
from .frame import FrameBuffer

_bound = False


def _lib() -> ctypes.CDLL:
    global _bound
    L = io.lib()
    if not _bound:
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.psim_presets_new.restype = vp
        L.psim_presets_new.argtypes = []
        L.psim_presets_destroy.restype = None
        L.psim_presets_destroy.argtypes = [vp]
        L.psim_presets_len.restype = sz
        L.psim_presets_len.argtypes = [vp]
        L.psim_presets_add_from_frame.restype = ctypes.c_long
        L.psim_presets_add_from_frame.argtypes = [vp, ctypes.c_char_p, vp]
        L.psim_presets_change_from_frame.restype = ctypes.c_int
        L.psim_presets_change_from_frame.argtypes = [vp, sz, ctypes.c_char_p, vp]
        L.psim_presets_duplicate.restype = ctypes.c_long
        L.psim_presets_duplicate.argtypes = [vp, sz, ctypes.c_char_p]
        L.psim_presets_delete.restype = ctypes.c_int
        L.psim_presets_delete.argtypes = [vp, sz]
        L.psim_preset_name.restype = ctypes.c_char_p
        L.psim_preset_name.argtypes = [vp, sz]
        L.psim_preset_particle_count.restype = ctypes.c_uint32
        L.psim_preset_particle_count.argtypes = [vp, sz]
        L.psim_preset_to_frame.restype = ctypes.c_int
        L.psim_preset_to_frame.argtypes = [vp, sz, vp, ctypes.c_uint32]
        _bound = True
    return L


class Preset:
    """One entry of a `Presets` list (a view: the data lives in the C library)."""

    def __init__(self, owner: "Presets", ind: int):
        self._owner, self._ind = owner, ind

    @property
    def name(self) -> str:
        return _lib().psim_preset_name(self._owner._h, self._ind).decode()

    @property
    def particle_count(self) -> int:
        return int(_lib().psim_preset_particle_count(self._owner._h, self._ind))

    def to_frame(self) -> FrameBuffer:
        """presets.rs:92-105: a new frame with default metadata, this preset's box and species, all its particles."""
        fb = FrameBuffer(max(self.particle_count, 1))
        if _lib().psim_preset_to_frame(self._owner._h, self._ind, fb.ptr, fb.capacity) != 0:
            raise IndexError(self._ind)
        return fb


class Presets:
    """presets.rs:122-154."""

    def __init__(self):
        self._h = ctypes.c_void_p(_lib().psim_presets_new())

    def close(self) -> None:
        if self._h:
            _lib().psim_presets_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_presets_len(self) -> int:
        return int(_lib().psim_presets_len(self._h))

    def get_preset(self, ind: int) -> Preset:
        if not 0 <= ind < self.get_presets_len():
            raise IndexError(ind)  # the reference indexes the Vec and panics
        return Preset(self, ind)

    def add_preset(self, name: str, frame: FrameBuffer) -> int:
        """add_preset(Preset::from_frame(name, frame)); returns the new preset's index."""
        ind = int(_lib().psim_presets_add_from_frame(self._h, name.encode(), frame.ptr))
        if ind < 0:
            raise ValueError("psim_presets_add_from_frame failed")
        return ind

    def duplicate_preset(self, ind: int, new_name: str) -> int:
        out = int(_lib().psim_presets_duplicate(self._h, ind, new_name.encode()))
        if out < 0:
            raise IndexError(ind)
        return out

    def delete_preset(self, ind: int) -> None:
        if _lib().psim_presets_delete(self._h, ind) != 0:
            raise IndexError(ind)

    def change_preset(self, name: str, frame: FrameBuffer, ind: int) -> None:
        """change_preset(Preset::from_frame(name, frame), ind): an index past the end is ignored (presets.rs:147-152)."""
        if _lib().psim_presets_change_from_frame(self._h, ind, name.encode(), frame.ptr) != 0:
            raise ValueError("psim_presets_change_from_frame failed")
