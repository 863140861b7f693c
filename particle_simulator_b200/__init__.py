"""B200-native particle stepper: drop-in for the per-timestep update of otcova/particle-simulator.

The product is two C-ABI shared libraries (see include/*.h):
  libparticle_io_c.so  the reference's particle_io C API (frame format, reader/writer/TCP) + scenes
  libpsim_b200.so      the CUDA (sm_100a) stepper: binning, 3x3-cell Mie force, leapfrog
This package only holds their sources (csrc/), the build recipe and thin ctypes mirrors of the
reference-facing interfaces for tests and benchmarks.
"""
from .frame import (HEADER_DTYPE, METADATA_DTYPE, PARTICLE_DTYPE, FrameBuffer, default_metadata,  # noqa: F401
                    packet_size)

__all__ = ["FrameBuffer", "default_metadata", "packet_size", "PARTICLE_DTYPE", "METADATA_DTYPE", "HEADER_DTYPE"]
