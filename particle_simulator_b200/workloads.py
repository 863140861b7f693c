"""The synthetic workloads BASELINE.json names, as particle_io frames.

All keep the reference's cell width (50 nm / 64 = 7.8125e-10 m, kernel.cuh:15-18 with
particle.rs:141-142) so that particles per cell match what the reference was designed for
(hex lattice at r0: 4.4 per cell, capacity 16).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import io
from .frame import FrameBuffer

CELL_WIDTH = 50e-9 / 64


@dataclass
class Workload:
    name: str
    description: str
    grid_log2: tuple[int, int]
    frame: FrameBuffer

    @property
    def particles(self) -> int:
        return self.frame.count


def _frame(capacity: int, grid_log2: tuple[int, int], storage: np.ndarray | None) -> FrameBuffer:
    fb = FrameBuffer(capacity, storage=storage)
    fb.metadata["box_width"] = CELL_WIDTH * (1 << grid_log2[0])
    fb.metadata["box_height"] = CELL_WIDTH * (1 << grid_log2[1])
    return fb


def lattice(n_side_x: int, n_side_y: int, grid_log2: tuple[int, int], distance_factor: float = 1.0,
            v_min: float = 1.0, v_max: float = 10.0, seed: int = 3, storage: np.ndarray | None = None,
            name: str = "lattice") -> Workload:
    """Perfect hex lattice (ParticleLattice::hex_square, presets.rs:16-46) centred in the box, speeds as the
    editor's default 1..10 m/s (editor.rs:178-182)."""
    fb = _frame(n_side_x * n_side_y, grid_log2, storage)
    cx, cy = float(fb.metadata["box_width"]) / 2, float(fb.metadata["box_height"]) / 2
    io.scene_hex_square(fb, n_side_x, n_side_y, (cx, cy), distance_factor, v_min, v_max, 0, seed)
    desc = (f"{n_side_x}x{n_side_y} hex lattice at {distance_factor:g} r0, speeds {v_min:g}-{v_max:g} m/s, "
            f"{1 << grid_log2[0]}x{1 << grid_log2[1]} cells, box {float(fb.metadata['box_width']) * 1e6:.2f}x"
            f"{float(fb.metadata['box_height']) * 1e6:.2f} um")
    return Workload(name, desc, grid_log2, fb)


def config_10m_solid(storage: np.ndarray | None = None) -> Workload:
    """BASELINE.json configs[2] (the one the metric is quoted on): 10M-particle solid lattice,
    2048x2048 cells, box 1.6 um."""
    return lattice(3162, 3163, (11, 11), 1.0, 1.0, 10.0, seed=3, storage=storage, name="10M-solid-lattice")


def config_1m_liquid(storage: np.ndarray | None = None) -> Workload:
    """BASELINE.json configs[1]: 1M-particle liquid-density box, 1024x1024 cells... the lattice at 1.05 r0 with
    150-250 m/s melts within a few hundred steps."""
    return lattice(1000, 1000, (10, 10), 1.05, 150.0, 250.0, seed=1, storage=storage, name="1M-liquid")


def slab_lattice(rank: int, nranks: int, rows_per_rank_log2: int = 11, storage: np.ndarray | None = None) -> Workload:
    """Weak-scaling workload: every rank owns a 2048-row slab holding its own 10M-particle lattice; the global
    grid is 2048 x (2048 * nranks) cells."""
    raise NotImplementedError
