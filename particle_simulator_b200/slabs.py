"""Host-side helpers of the slab decomposition (one process per GPU).

The domain is cut into `world` slabs of cell rows: equal numbers of rows (rank r owns rows
[r, r + 1) * (cells in y / world)), or rows [bounds[r], bounds[r + 1]) of a list of world + 1 boundaries
(stepper.balance_rows cuts a scene so that every slab holds about the same number of particles). Everything here is plain numpy / torch.distributed plumbing: which
slab a record belongs to, handing the NCCL unique id around, and reductions of timings over ranks.
"""
from __future__ import annotations

import numpy as np


def slab_rows(rank: int, world: int, grid_y_log2: int) -> tuple[int, int]:
    """(first owned cell row, number of owned rows) of slab `rank`."""
    rows = 1 << grid_y_log2
    if world < 1 or rows % world or rows // world < 2 or not 0 <= rank < world:
        raise ValueError(f"{rows} cell rows do not split into {world} slabs of at least 2 rows")
    per = rows // world
    return rank * per, per


def slab_of(y: np.ndarray, world: int, grid_y_log2: int, bounds=None) -> np.ndarray:
    """Owner slab of fixed-point y coordinates: cell row = y >> (32 - LY), kernel.cuh:225."""
    row = (np.asarray(y, dtype=np.uint32) >> np.uint32(32 - grid_y_log2)).astype(np.int64)
    if bounds is not None:
        assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == 1 << grid_y_log2
        return np.searchsorted(np.asarray(bounds[1:-1], dtype=np.int64), row, side="right")
    return row // slab_rows(0, world, grid_y_log2)[1]


def split_by_slab(particles: np.ndarray, world: int, grid_y_log2: int, bounds=None) -> list[np.ndarray]:
    """The records of each slab, in input order (what each rank's ingest keeps of a whole scene)."""
    owner = slab_of(particles["y"], world, grid_y_log2, bounds)
    live = particles["ty"] >= 0
    return [particles[live & (owner == r)] for r in range(world)]


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Every rank gets `payload` of rank `src` (the 128-byte NCCL unique id)."""
    import torch

    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def reduce_scalar(dist, value: float, op: str = "max", device=None) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())
