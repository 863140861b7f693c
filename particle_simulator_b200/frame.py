"""numpy views of the `particle_io` wire format (include/particle_io.h).

Layouts follow the reference's `#[repr(C)]` structs (particle_io/src/particle.rs:10-18, 33-41,
111-130, 192-204): Particle 20 B, FrameMetadata 80 B, FrameHeader 96 B, little-endian.
"""
from __future__ import annotations

import ctypes

import numpy as np

PARTICLE_DTYPE = np.dtype(
    [("x", "<u4"), ("y", "<u4"), ("vx", "<f4"), ("vy", "<f4"), ("ty", "<i4")]
)
MIE_DTYPE = np.dtype([("sigma", "<f4"), ("epsilon", "<f4"), ("n", "<f4"), ("m", "<f4")])
METADATA_DTYPE = np.dtype(
    [
        ("particles", MIE_DTYPE, (2,)),
        ("cursor_pos", "<f4", (2,)),
        ("cursor_size", "<f4"),
        ("step_dt", "<f4"),
        ("steps_per_frame", "<u4"),
        ("box_width", "<f4"),
        ("box_height", "<f4"),
        ("data_structure", "<u4"),
        ("device", "<u4"),
        ("gpu_threads_per_block_log2", "<u4"),
        ("_padding", "<u4", (2,)),
    ]
)
HEADER_DTYPE = np.dtype(
    [
        ("signature_start", "u1", (4,)),
        ("particle_count", "<u4"),
        ("metadata", METADATA_DTYPE),
        ("signature_end", "u1", (4,)),
        ("_padding", "<u4"),
    ]
)
assert PARTICLE_DTYPE.itemsize == 20
assert METADATA_DTYPE.itemsize == 80
assert HEADER_DTYPE.itemsize == 96

SIGNATURE_START = bytes([0x36, 0xBC, 0xE9, 0xBD])  # particle.rs:207
SIGNATURE_END = bytes([0xAC, 0xC4, 0x12, 0xEC])  # particle.rs:208

# particle.rs:52-57, 80-86
COMPACT_ARRAY, MATRIX_BUCKETS = 0, 1
DEVICE_GPU, DEVICE_CPU_THREAD_POOL, DEVICE_CPU_MAIN_THREAD = 0, 1, 2

K_B = np.float32(1.380649e-23)
PARTICLE_MASS = np.float32(6.63352599e-26)  # cuda_simulator/src/particle.cuh:51


def packet_size(particle_count: int) -> int:
    """particle.rs:225-227"""
    return HEADER_DTYPE.itemsize + PARTICLE_DTYPE.itemsize * int(particle_count)


def default_metadata() -> np.ndarray:
    """FrameMetadata::default(), particle.rs:132-165 (a 0-d structured array)."""
    m = np.zeros((), dtype=METADATA_DTYPE)
    m["cursor_pos"] = (-1.0, -1.0)
    m["cursor_size"] = 0.05
    m["step_dt"] = 50e-15
    m["steps_per_frame"] = 100
    m["box_width"] = 50e-9
    m["box_height"] = 50e-9
    m["data_structure"] = MATRIX_BUCKETS
    m["device"] = DEVICE_GPU
    m["gpu_threads_per_block_log2"] = 7
    m["particles"][0] = (3.609e-10, np.float32(105.79) * K_B, 14.08, 6.0)
    m["particles"][1] = (3.404e-10, np.float32(117.84) * K_B, 12.085, 6.0)
    return m


class FrameBuffer:
    """A caller-owned frame: 96-byte header followed by `capacity` particle records.

    This is the memory shape every C function taking `FrameHeader*` expects
    (particle_io/c_api/src/particle.rs:43-58).
    """

    def __init__(self, capacity: int, metadata: np.ndarray | None = None, storage: np.ndarray | None = None):
        """`storage`: optional uint8 array of at least packet_size(capacity) bytes to build the frame
        in (e.g. a view of page-locked memory); by default the frame owns ordinary host memory."""
        self.capacity = int(capacity)
        if storage is None:
            self.raw = np.zeros(packet_size(self.capacity), dtype=np.uint8)
        else:
            if storage.dtype != np.uint8 or storage.size < packet_size(self.capacity) or storage.ctypes.data % 4:
                raise ValueError("storage must be an aligned uint8 array of at least packet_size(capacity) bytes")
            self.raw = storage[: packet_size(self.capacity)]
            self.raw[: HEADER_DTYPE.itemsize] = 0
        self.header = self.raw[: HEADER_DTYPE.itemsize].view(HEADER_DTYPE)[0:1]
        self._all = self.raw[HEADER_DTYPE.itemsize :].view(PARTICLE_DTYPE)
        self.header["signature_start"] = np.frombuffer(SIGNATURE_START, dtype=np.uint8)
        self.header["signature_end"] = np.frombuffer(SIGNATURE_END, dtype=np.uint8)
        self.header["metadata"] = default_metadata() if metadata is None else metadata
        self.header["particle_count"] = 0

    # -- header fields -----------------------------------------------------------------------
    @property
    def count(self) -> int:
        return int(self.header["particle_count"][0])

    @count.setter
    def count(self, n: int) -> None:
        self.header["particle_count"] = n

    @property
    def metadata(self) -> np.ndarray:
        """Writable view of the embedded FrameMetadata."""
        return self.header["metadata"][0:1].reshape(())

    @property
    def particles(self) -> np.ndarray:
        """The first `count` records."""
        return self._all[: self.count]

    @property
    def all_particles(self) -> np.ndarray:
        return self._all

    @property
    def ptr(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(self.raw.ctypes.data)

    def is_valid(self) -> bool:
        return (
            bytes(self.header["signature_start"][0]) == SIGNATURE_START
            and bytes(self.header["signature_end"][0]) == SIGNATURE_END
        )

    def tobytes(self) -> bytes:
        """The bytes that go on the wire: header + `count` records."""
        return self.raw[: packet_size(self.count)].tobytes()

    def set_particles(self, particles: np.ndarray) -> None:
        n = len(particles)
        if n > self.capacity:
            raise ValueError(f"{n} particles do not fit a frame of capacity {self.capacity}")
        self._all[:n] = particles
        self.count = n

    @classmethod
    def from_bytes(cls, data: bytes) -> "FrameBuffer":
        header = np.frombuffer(data[: HEADER_DTYPE.itemsize], dtype=HEADER_DTYPE)[0]
        n = int(header["particle_count"])
        if len(data) != packet_size(n):
            raise ValueError("byte length does not match particle_count")
        fb = cls(n)
        fb.raw[:] = np.frombuffer(data, dtype=np.uint8)
        return fb

    def copy(self, capacity: int | None = None) -> "FrameBuffer":
        fb = FrameBuffer(self.capacity if capacity is None else capacity, self.metadata)
        fb.set_particles(self.particles)
        return fb


def force0_r(mie: np.ndarray) -> float:
    """r0 = sigma (n/m)^(1/(n-m)) in double (particle.rs:44-49)."""
    n, m, sigma = float(mie["n"]), float(mie["m"]), float(mie["sigma"])
    return sigma * (n / m) ** (1.0 / (n - m))
