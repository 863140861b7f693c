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


def rebalance_across_ranks(dist, st, make_stepper, device=None):
    """Move the slab boundaries of a running one-process-per-GPU decomposition to where the particles are now
    (BASELINE.json configs[4]). Collective over `dist`, called between two frames:

      1. every rank snapshots and downloads its own particles and histograms them by global cell row;
      2. the histograms are summed over the ranks (one all-reduce) and every rank cuts the same boundaries
         (psim_balance_rows_hist: equal shares of the particles, every slab at least 2 rows);
      3. every rank sends each of its particles to the rank that owns its row now (all-to-all over the records,
         20 bytes each; slabs are bands of rows, so almost everything stays where it is);
      4. the stepper is re-created on the new rows (`make_stepper(bounds)` must return a new, communicator-initialised
         Stepper: the slab geometry is fixed at psim_create) and takes the records in as a scene.

    Like SlabGroup.rebalance this restarts the step / re-bin schedule (an upload's worth of work, for when the imbalance
    has grown). The received records arrive in ascending source rank = ascending row order, each rank's in its sorted
    order, so the ingest's stable sort reproduces the cell-sorted order a single slab would have.
    Returns (new stepper, boundaries)."""
    import torch

    from .frame import PARTICLE_DTYPE, FrameBuffer
    from .stepper import balance_rows_hist

    world, rank = dist.get_world_size(), dist.get_rank()
    ly = st.grid_log2[1]
    meta = st.get_metadata()
    st.sync()
    st.snapshot_async()
    mine = st.download().particles
    rows = (mine["y"] >> np.uint32(32 - ly)).astype(np.int64)
    hist = torch.from_numpy(np.bincount(rows, minlength=1 << ly).astype(np.int64)).to(device)
    dist.all_reduce(hist)
    bounds = balance_rows_hist(hist.cpu().numpy().astype(np.uint64), ly, world)
    owner = np.searchsorted(np.asarray(bounds[1:-1], dtype=np.int64), rows, side="right")
    order = np.argsort(owner, kind="stable")
    send_counts = np.bincount(owner, minlength=world).astype(np.int64)
    send = torch.from_numpy(np.ascontiguousarray(mine[order]).view(np.uint8).reshape(-1).copy()).to(device)
    counts_dev = torch.from_numpy(send_counts).to(device)
    recv_counts_dev = torch.zeros_like(counts_dev)
    dist.all_to_all_single(recv_counts_dev, counts_dev)
    recv_counts = recv_counts_dev.cpu().numpy()
    recv = torch.empty(int(recv_counts.sum()) * 20, dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=(recv_counts * 20).tolist(),
                           input_split_sizes=(send_counts * 20).tolist())
    records = recv.cpu().numpy().view(PARTICLE_DTYPE)
    st.close()
    new = make_stepper(bounds)
    fb = FrameBuffer(max(len(records), 1), meta)
    fb.set_particles(records)
    new.upload(fb)
    return new, bounds
