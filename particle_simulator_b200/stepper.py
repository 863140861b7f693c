"""ctypes binding of libpsim_b200.so (include/psim_b200.h): the B200 particle stepper.

`Stepper` mirrors the reference's `Kernel` object (cuda_simulator/src/kernel.cuh:22-132) the way its
main loop uses it (cuda_simulator/src/cuda_simulator.cu:7-38):

    reference                                   here
    kernel_prepare_frame + kernel.write(id)     Stepper.upload(frame)
    kernel.write_metadata(id)                   Stepper.set_metadata(meta)
    kernel.run_async(src, dst)                  Stepper.run_frame_async()
    kernel.sync()                               Stepper.sync()
    kernel.read(id) + frame_compact             Stepper.download(frame)

There is no CPU fallback: if the CUDA library is missing or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _build
from .frame import METADATA_DTYPE, FrameBuffer

SCHEDULE_REFERENCE = 0
SCHEDULE_NATIVE = 1

_lib = None


class PsimError(RuntimeError):
    pass


class CConfig(ctypes.Structure):
    _fields_ = [
        ("grid_x_log2", ctypes.c_uint32),
        ("grid_y_log2", ctypes.c_uint32),
        ("max_particles", ctypes.c_uint32),
        ("schedule", ctypes.c_uint32),
        ("rebin_every", ctypes.c_uint32),
        ("device", ctypes.c_int32),
        ("use_graph", ctypes.c_uint32),
        ("slab_rank", ctypes.c_uint32),
        ("slab_count", ctypes.c_uint32),
        ("ghost_capacity", ctypes.c_uint32),
        ("migrant_capacity", ctypes.c_uint32),
        ("ingest_capacity", ctypes.c_uint32),
        ("snapshot_buffers", ctypes.c_uint32),
        ("slab_bounds", ctypes.c_uint32 * 4),
        ("species_physics", ctypes.c_uint32),
    ]


class CSlabInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "slab_rank", "slab_count", "first_row", "rows", "first_local_row", "local_rows", "particles",
        "ghost_below", "ghost_above", "ghost_capacity", "migrant_capacity")] + [("_reserved", ctypes.c_uint32 * 5)]


class CTileStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("float_path", "tiles", "tiles_staged", "max_columns",
                                               "max_row_particles", "_reserved")] + \
               [("threads_live", ctypes.c_uint64), ("threads_launched", ctypes.c_uint64)]


def lib() -> ctypes.CDLL:
    """Load libpsim_b200.so; raises if it has not been built (the product has no other path)."""
    global _lib
    if _lib is None:
        path = os.environ.get("PSIM_LIB") or _build.LIB_PSIM  # PSIM_LIB: another build of the same library (A/B runs)
        if not os.path.exists(path):
            raise PsimError(f"{path} is missing: run `python -m particle_simulator_b200._build` "
                            "(there is no CPU fallback for the stepper)")
        L = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        L.psim_default_config.restype = CConfig
        L.psim_default_config.argtypes = []
        L.psim_create.restype = ctypes.c_int
        L.psim_create.argtypes = [ctypes.POINTER(CConfig), ctypes.POINTER(vp)]
        L.psim_destroy.restype = None
        L.psim_destroy.argtypes = [vp]
        L.psim_slab_bounds_of.restype = None
        L.psim_slab_bounds_of.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.POINTER(ctypes.c_uint32)]
        L.psim_last_error.restype = ctypes.c_char_p
        L.psim_last_error.argtypes = [vp]
        for name, args in {
            "psim_set_stream": [vp, vp],
            "psim_upload_frame": [vp, vp],
            "psim_upload_device": [vp, vp, vp, ctypes.c_uint32],
            "psim_set_metadata": [vp, vp],
            "psim_get_metadata": [vp, vp],
            "psim_run_frame_async": [vp],
            "psim_step_async": [vp, ctypes.c_uint32],
            "psim_rebin_async": [vp],
            "psim_snapshot_async": [vp],
            "psim_set_snapshot_stride": [vp, ctypes.c_uint32],
            "psim_sync": [vp],
            "psim_download_frame": [vp, vp],
            "psim_download_frame_ex": [vp, ctypes.c_uint32, vp],
            "psim_stage_frame_async": [vp, vp],
            "psim_upload_staged": [vp],
            "psim_download_frame_begin": [vp, ctypes.c_uint32, vp],
            "psim_download_frame_end": [vp],
            "psim_get_cell_start": [vp, vp],
            "psim_enable_step_timing": [vp, ctypes.c_int],
            "psim_get_step_timing": [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)],
            "psim_device_state": [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)],
            "psim_slab_info": [vp, ctypes.POINTER(CSlabInfo)],
            "psim_tile_stats": [vp, ctypes.POINTER(CTileStats)],
            "psim_comm_unique_id": [vp],
            "psim_comm_init": [vp, vp],
            "psim_group_create": [ctypes.POINTER(vp), ctypes.c_uint32, ctypes.POINTER(vp)],
            "psim_balance_rows": [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)],
            "psim_balance_rows_hist": [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)],
            "psim_group_upload_frame": [vp, vp],
            "psim_group_set_metadata": [vp, vp],
            "psim_group_run_frame_async": [vp],
            "psim_group_step_async": [vp, ctypes.c_uint32],
            "psim_group_rebin_async": [vp],
            "psim_group_snapshot_async": [vp],
            "psim_group_sync": [vp],
            "psim_group_download_frame": [vp, vp],
        }.items():
            if os.environ.get("PSIM_LIB") and not hasattr(L, name):
                continue  # an older build of the library under A/B test (tools/ab_step.py) lacks the newest entry points
            getattr(L, name).restype = ctypes.c_int
            getattr(L, name).argtypes = args
        L.psim_group_destroy.restype = None
        L.psim_group_destroy.argtypes = [vp]
        L.psim_group_last_error.restype = ctypes.c_char_p
        L.psim_group_last_error.argtypes = [vp]
        L.psim_group_particle_count.restype = ctypes.c_uint32
        L.psim_group_particle_count.argtypes = [vp]
        L.psim_halo_mode.restype = ctypes.c_int
        L.psim_halo_mode.argtypes = [vp]
        L.psim_particle_count.restype = ctypes.c_uint32
        L.psim_cell_count.restype = ctypes.c_uint32
        for name in ("psim_steps_executed", "psim_rebins_executed", "psim_kernel_launches", "psim_migrants_sent"):
            if name == "psim_migrants_sent" and os.environ.get("PSIM_LIB") and not hasattr(L, name):
                continue  # an older build of the library under A/B test (tools/ab_step.py)
            getattr(L, name).restype = ctypes.c_uint64
            getattr(L, name).argtypes = [vp]
        for name in ("psim_particle_count", "psim_cell_count"):
            getattr(L, name).argtypes = [vp]
        _lib = L
    return _lib


class Stepper:
    def __init__(self, grid_log2: tuple[int, int] = (6, 6), max_particles: int = 65536,
                 schedule: int = SCHEDULE_REFERENCE, rebin_every: int = 0, device: int = -1,
                 use_graph: bool = False, slab_rank: int = 0, slab_count: int = 1, ghost_capacity: int = 0,
                 species_physics: bool = False,
                 migrant_capacity: int = 0, ingest_capacity: int = 0, snapshot_buffers: int = 1,
                 bounds=None):
        """`bounds`: slab_count + 1 cell-row boundaries shared by all slabs (balance_rows); None: equal shares."""
        L = lib()
        cfg = L.psim_default_config()
        cfg.grid_x_log2, cfg.grid_y_log2 = grid_log2
        cfg.max_particles = max_particles
        cfg.schedule = schedule
        cfg.rebin_every = rebin_every
        cfg.device = device
        cfg.use_graph = 1 if use_graph else 0
        cfg.species_physics = 1 if species_physics else 0
        cfg.slab_rank, cfg.slab_count = slab_rank, slab_count
        cfg.ghost_capacity, cfg.migrant_capacity, cfg.ingest_capacity = ghost_capacity, migrant_capacity, ingest_capacity
        cfg.snapshot_buffers = snapshot_buffers
        if bounds is not None:
            b = [int(v) for v in bounds]
            if len(b) != slab_count + 1:
                raise ValueError(f"{slab_count} slabs have {slab_count + 1} row boundaries, got {len(b)}")
            arr = (ctypes.c_uint32 * (slab_count + 1))(*b)
            L.psim_slab_bounds_of(arr, slab_rank, slab_count, cfg.slab_bounds)
        self._h = ctypes.c_void_p()
        rc = L.psim_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            self._h = ctypes.c_void_p()
            raise PsimError(f"psim_create failed ({rc}): {L.psim_last_error(None).decode()}")
        self.grid_log2 = tuple(grid_log2)
        self.max_particles = max_particles

    # -- plumbing ----------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PsimError(f"psim error {rc}: {lib().psim_last_error(self._h).decode()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().psim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the Kernel-shaped API ---------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(lib().psim_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def upload(self, frame: FrameBuffer) -> None:
        self._check(lib().psim_upload_frame(self._h, frame.ptr))

    def upload_device(self, metadata: np.ndarray, d_particles: int, count: int) -> None:
        meta = np.ascontiguousarray(metadata, dtype=METADATA_DTYPE)
        self._check(lib().psim_upload_device(self._h, ctypes.c_void_p(meta.ctypes.data),
                                             ctypes.c_void_p(d_particles), count))

    def set_metadata(self, metadata: np.ndarray) -> None:
        meta = np.ascontiguousarray(metadata, dtype=METADATA_DTYPE)
        self._check(lib().psim_set_metadata(self._h, ctypes.c_void_p(meta.ctypes.data)))

    def get_metadata(self) -> np.ndarray:
        meta = np.zeros((), dtype=METADATA_DTYPE)
        self._check(lib().psim_get_metadata(self._h, ctypes.c_void_p(meta.ctypes.data)))
        return meta

    def run_frame_async(self) -> None:
        self._check(lib().psim_run_frame_async(self._h))

    def step_async(self, steps: int = 1) -> None:
        self._check(lib().psim_step_async(self._h, steps))

    def rebin_async(self) -> None:
        self._check(lib().psim_rebin_async(self._h))

    def snapshot_async(self) -> None:
        self._check(lib().psim_snapshot_async(self._h))

    def set_snapshot_stride(self, stride: int) -> None:
        """Snapshots hold every `stride`-th particle from now on (decimated frames for display)."""
        self._check(lib().psim_set_snapshot_stride(self._h, stride))

    def sync(self) -> None:
        self._check(lib().psim_sync(self._h))

    def download(self, frame: FrameBuffer | None = None, age: int = 0) -> FrameBuffer:
        """The snapshot packed `age` snapshots ago (0: the latest; 1 needs snapshot_buffers=2)."""
        if frame is None:
            frame = FrameBuffer(max(self.particle_count, 1))
        frame.count = frame.capacity
        self._check(lib().psim_download_frame_ex(self._h, age, frame.ptr))
        return frame

    # -- pipelined frames: copies on their own streams while frames compute --------------------
    def stage_async(self, frame: FrameBuffer) -> None:
        """Start the host-to-device copy of `frame`; keep it untouched until upload_staged() has returned."""
        self._check(lib().psim_stage_frame_async(self._h, frame.ptr))

    def upload_staged(self) -> None:
        self._check(lib().psim_upload_staged(self._h))

    def download_begin(self, frame: FrameBuffer, age: int = 0) -> None:
        frame.count = frame.capacity
        self._check(lib().psim_download_frame_begin(self._h, age, frame.ptr))

    def download_end(self) -> None:
        self._check(lib().psim_download_frame_end(self._h))

    # -- slab decomposition, one process per slab (NCCL) -------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId (rank 0 calls it and distributes the 128 bytes to the other ranks)."""
        buf = ctypes.create_string_buffer(128)
        rc = lib().psim_comm_unique_id(buf)
        if rc != 0:
            raise PsimError(f"psim_comm_unique_id failed ({rc}): {lib().psim_last_error(None).decode()}")
        return buf.raw

    def comm_init(self, unique_id: bytes) -> None:
        assert len(unique_id) == 128
        self._check(lib().psim_comm_init(self._h, ctypes.c_char_p(unique_id)))

    @property
    def halo_mode(self) -> int:
        """0: single slab, 1: send/recv after every step, 2: pushed by the step kernel over peer memory."""
        return int(lib().psim_halo_mode(self._h))

    def slab_info(self) -> dict:
        info = CSlabInfo()
        self._check(lib().psim_slab_info(self._h, ctypes.byref(info)))
        return {n: int(getattr(info, n)) for n, _ in CSlabInfo._fields_ if n != "_reserved"}

    def tile_stats(self) -> dict:
        st = CTileStats()
        self._check(lib().psim_tile_stats(self._h, ctypes.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in CTileStats._fields_ if n != "_reserved"}

    # -- introspection -----------------------------------------------------------------------
    @property
    def particle_count(self) -> int:
        return int(lib().psim_particle_count(self._h))

    @property
    def cell_count(self) -> int:
        return int(lib().psim_cell_count(self._h))

    @property
    def steps_executed(self) -> int:
        return int(lib().psim_steps_executed(self._h))

    @property
    def rebins_executed(self) -> int:
        return int(lib().psim_rebins_executed(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().psim_kernel_launches(self._h))

    @property
    def migrants_sent(self) -> int:
        """Particles this slab has handed to its neighbours at re-bins so far."""
        return int(lib().psim_migrants_sent(self._h))

    def cell_start(self) -> np.ndarray:
        out = np.zeros(self.cell_count + 1, dtype=np.uint32)
        self._check(lib().psim_get_cell_start(self._h, ctypes.c_void_p(out.ctypes.data)))
        return out

    def enable_step_timing(self, enable: bool = True) -> None:
        self._check(lib().psim_enable_step_timing(self._h, 1 if enable else 0))

    def step_timing(self) -> tuple[float, int]:
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        self._check(lib().psim_get_step_timing(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value


def balance_rows(frame: FrameBuffer, grid_y_log2: int, slab_count: int) -> list[int]:
    """psim_balance_rows: slab_count + 1 cell-row boundaries that give every slab about the same number of the
    frame's live particles (host code; needs no GPU)."""
    out = (ctypes.c_uint32 * (slab_count + 1))()
    rc = lib().psim_balance_rows(frame.ptr, grid_y_log2, slab_count, out)
    if rc != 0:
        raise PsimError(f"psim_balance_rows failed ({rc}): {lib().psim_last_error(None).decode()}")
    return list(out)


def balance_rows_hist(row_counts: np.ndarray, grid_y_log2: int, slab_count: int) -> list[int]:
    """psim_balance_rows_hist: the same cut from a histogram of live particles per global cell row."""
    hist = np.ascontiguousarray(row_counts, dtype=np.uint64)
    assert hist.size == 1 << grid_y_log2
    out = (ctypes.c_uint32 * (slab_count + 1))()
    rc = lib().psim_balance_rows_hist(ctypes.c_void_p(hist.ctypes.data), grid_y_log2, slab_count, out)
    if rc != 0:
        raise PsimError(f"psim_balance_rows_hist failed ({rc}): {lib().psim_last_error(None).decode()}")
    return list(out)


class SlabGroup:
    """All slabs of a row decomposition inside one process on one device (psim_group_*): the same
    slabs, halo exchange and migration as the one-process-per-GPU NCCL mode, with device-to-device
    copies as the transport. Results are bit-identical to a single `Stepper` on the same scene."""

    def __init__(self, grid_log2: tuple[int, int], slab_count: int, max_particles_per_slab: int,
                 ingest_capacity: int = 0, schedule: int = SCHEDULE_REFERENCE, rebin_every: int = 0, device: int = -1,
                 ghost_capacity: int = 0, migrant_capacity: int = 0, bounds=None):
        """`bounds`: slab_count + 1 cell-row boundaries (balance_rows(scene, ...)); None: equal numbers of rows."""
        self.grid_log2, self.bounds = grid_log2, None if bounds is None else [int(b) for b in bounds]
        self._make = lambda r, bnds, cap: Stepper(grid_log2, cap, schedule, rebin_every, device, slab_rank=r,
                                                  slab_count=slab_count, ghost_capacity=ghost_capacity,
                                                  migrant_capacity=migrant_capacity, ingest_capacity=ingest_capacity,
                                                  bounds=bnds)
        self._build(slab_count, self.bounds, max_particles_per_slab)

    def _build(self, slab_count: int, bounds, max_particles_per_slab: int) -> None:
        self.max_particles_per_slab = max_particles_per_slab
        self.slabs = [self._make(r, bounds, max_particles_per_slab) for r in range(slab_count)]
        arr = (ctypes.c_void_p * slab_count)(*[s._h for s in self.slabs])
        self._g = ctypes.c_void_p()
        rc = lib().psim_group_create(arr, slab_count, ctypes.byref(self._g))
        if rc != 0:
            raise PsimError(f"psim_group_create failed ({rc}): {lib().psim_last_error(None).decode()}")

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PsimError(f"psim error {rc}: {lib().psim_group_last_error(self._g).decode()}")

    def close(self) -> None:
        if getattr(self, "_g", None) and self._g.value:
            lib().psim_group_destroy(self._g)
            self._g = ctypes.c_void_p()
        for s in getattr(self, "slabs", []):
            s.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, frame: FrameBuffer) -> None:
        self._check(lib().psim_group_upload_frame(self._g, frame.ptr))

    def rebalance(self, max_particles_per_slab: int = 0) -> list[int]:
        """Move the slab boundaries to where the particles are now: the latest state is
        downloaded, the rows are cut again by psim_balance_rows, the slabs are re-created on the new rows and the
        state goes back in -- an upload's worth of work, for between frames when the imbalance has grown. The
        step / re-bin countdown restarts as after any upload. Returns the new boundaries."""
        self.snapshot_async()
        frame = self.download()
        bounds = balance_rows(frame, self.grid_log2[1], len(self.slabs))
        cap = max_particles_per_slab or self.max_particles_per_slab
        count = len(self.slabs)
        self.close()
        self.bounds = bounds
        self._build(count, bounds, cap)
        self.upload(frame)
        return bounds

    def set_metadata(self, metadata: np.ndarray) -> None:
        meta = np.ascontiguousarray(metadata, dtype=METADATA_DTYPE)
        self._check(lib().psim_group_set_metadata(self._g, ctypes.c_void_p(meta.ctypes.data)))

    def run_frame_async(self) -> None:
        self._check(lib().psim_group_run_frame_async(self._g))

    def step_async(self, steps: int = 1) -> None:
        self._check(lib().psim_group_step_async(self._g, steps))

    def rebin_async(self) -> None:
        self._check(lib().psim_group_rebin_async(self._g))

    def snapshot_async(self) -> None:
        self._check(lib().psim_group_snapshot_async(self._g))

    def sync(self) -> None:
        self._check(lib().psim_group_sync(self._g))

    @property
    def particle_count(self) -> int:
        return int(lib().psim_group_particle_count(self._g))

    def download(self, frame: FrameBuffer | None = None) -> FrameBuffer:
        if frame is None:
            frame = FrameBuffer(max(self.particle_count, 1))
        frame.count = frame.capacity
        self._check(lib().psim_group_download_frame(self._g, frame.ptr))
        return frame
