"""The synthetic workloads BASELINE.json names, as particle_io frames.

All keep the reference's cell width (50 nm / 64 = 7.8125e-10 m, kernel.cuh:15-18 with
particle.rs:141-142) so that particles per cell match what the reference was designed for
(hex lattice at r0: 4.4 per cell, capacity 16).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import io
from .frame import FrameBuffer, default_metadata

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


def slab_crystal_geometry(world: int, per_slab: int = 3162 * 3163, rows_per_slab_log2: int = 11,
                          grid_x_log2: int = 11, margin_cells: int = 4) -> dict:
    """Weak-scaling workload (BASELINE.json: 1/2/4/8 B200): ONE hex crystal at r0 spanning `world` slabs of
    2048 cell rows each, `per_slab` particles per slab on average, so that every slab holds the same work and
    every slab boundary cuts through the crystal (real halo traffic, real migration). The box grows in y only:
    2048 x (2048 * world) cells of the reference's cell width."""
    assert world >= 1 and world & (world - 1) == 0, "slab counts are powers of two (the grid is)"
    meta = default_metadata()
    grid = (grid_x_log2, rows_per_slab_log2 + world.bit_length() - 1)
    width, height = CELL_WIDTH * (1 << grid[0]), CELL_WIDTH * (1 << grid[1])
    meta["box_width"], meta["box_height"] = width, height
    r0 = io.force0_r(meta)
    ry = float(np.sin(np.pi / 3)) * r0
    ny = int((height - 2 * margin_cells * CELL_WIDTH) / ry)
    nx = int(round(per_slab * world / ny))
    assert nx * r0 < width - 2 * margin_cells * CELL_WIDTH
    return {"grid_log2": grid, "box": (width, height), "nx": nx, "ny": ny, "r0": r0, "ry": ry,
            "start_y": height / 2 - ry * (ny - 1) / 2}


def slab_crystal(rank: int, world: int, storage_factory=None, per_slab: int = 3162 * 3163,
                 rows_per_slab_log2: int = 11, grid_x_log2: int = 11, seed: int = 3) -> Workload:
    """The lattice rows of slab_crystal_geometry() that fall into slab `rank` (plus one row either side: the
    stepper keeps the records of its own cell rows and skips the rest, kernel.cuh:222-226 semantics)."""
    geo = slab_crystal_geometry(world, per_slab, rows_per_slab_log2, grid_x_log2)
    width, height = geo["box"]
    slab_h = height / world
    lo = int(np.floor((rank * slab_h - geo["start_y"]) / geo["ry"])) - 1
    hi = int(np.ceil(((rank + 1) * slab_h - geo["start_y"]) / geo["ry"])) + 1
    lo, hi = max(lo, 0), min(hi, geo["ny"])
    count = geo["nx"] * (hi - lo)
    storage = storage_factory(count) if storage_factory else None
    fb = FrameBuffer(count, storage=storage)
    fb.metadata["box_width"], fb.metadata["box_height"] = width, height
    io.scene_hex_rows(fb, geo["nx"], geo["ny"], (lo, hi), (width / 2, height / 2), 1.0, 1.0, 10.0, 0, seed)
    desc = (f"one {geo['nx']}x{geo['ny']} hex crystal at 1 r0 across {world} slabs of {1 << rows_per_slab_log2} "
            f"cell rows, speeds 1-10 m/s, {1 << geo['grid_log2'][0]}x{1 << geo['grid_log2'][1]} cells, box "
            f"{width * 1e6:.2f}x{height * 1e6:.2f} um; slab {rank} is handed lattice rows [{lo}, {hi})")
    return Workload(f"10M-solid-lattice-per-slab-x{world}", desc, geo["grid_log2"], fb)


def clustered_mixed(grid_log2: tuple[int, int] = (10, 10), clusters: int = 6, side: int = 150, gas: int = 20000,
                    seed: int = 5, lopsided: bool = True) -> Workload:
    """BASELINE.json configs[4]: mixed-species clustered scene -- liquid-density droplets of species 0 and 1 in a mostly
    empty box plus a fast gas background of both species (heterogeneous density, load imbalance, high migration
    rate). Both species are stepped with species-0 parameters, as the reference does (kernel_bucket.cuh:52); `ty` is a
    label that travels with the particle. `lopsided` puts most droplets in the lower third of the box, so that a row
    decomposition is badly balanced."""
    rng = np.random.default_rng(seed)
    fb = _frame(clusters * side * side + gas, grid_log2, None)
    # a 300-600 m/s gas needs the shorter step: at the default 50 fs it crosses more than half a cell between two
    # re-bins and particles meet without ever having interacted (the reference blows up the same way)
    fb.metadata["step_dt"] = 10e-15
    w, h = float(fb.metadata["box_width"]), float(fb.metadata["box_height"])
    r0 = io.force0_r(fb.metadata)
    half = 0.5 * side * r0 * 1.05 + 4 * CELL_WIDTH
    centres: list[tuple[float, float]] = []
    attempts = 0
    while len(centres) < clusters:
        attempts += 1
        assert attempts < 10000, "the droplets do not fit the box"
        cx = rng.uniform(half, w - half)
        top = h / 3 if lopsided and len(centres) < clusters - 1 else h
        cy = rng.uniform(half, max(top, 2.2 * half) - half) if top > 2 * half else rng.uniform(half, h - half)
        if all(abs(cx - x) > 2 * half or abs(cy - y) > 2 * half for x, y in centres):
            centres.append((cx, cy))
    for k, c in enumerate(centres):
        io.scene_hex_square(fb, side, side, c, 1.05, 150.0, 250.0, k % 2, seed + 17 * k)
    if gas:
        io.scene_gas(fb, gas // 2, 3 * CELL_WIDTH, 3 * r0, 300.0, 600.0, 0, seed + 1000)
        io.scene_gas(fb, gas - gas // 2, 3 * CELL_WIDTH, 3 * r0, 300.0, 600.0, 1, seed + 1001)
    desc = (f"{clusters} droplets of {side}x{side} at 1.05 r0 (species alternate), {gas} gas particles at 300-600 m/s, "
            f"{1 << grid_log2[0]}x{1 << grid_log2[1]} cells")
    return Workload("mixed-species-clusters", desc, grid_log2, fb)


def heat(frame: FrameBuffer, factor: float) -> None:
    """The heating ramp of BASELINE.json configs[2]: every velocity times `factor`, on the host, between frames
    (the scene then goes back through the ordinary upload path; the reference has no thermostat)."""
    p = frame.particles
    p["vx"] *= np.float32(factor)
    p["vy"] *= np.float32(factor)
