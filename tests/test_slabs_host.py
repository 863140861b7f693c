"""Host logic of the N > 1 path on CPU: the slab partition of a scene, and the torch.distributed plumbing
(`gloo`, world_size 2) that bench.py and the NCCL worker use around the stepper."""
import os
import socket

import numpy as np
import pytest

from particle_simulator_b200 import FrameBuffer, io, slabs, workloads


def test_slab_rows_and_owner():
    assert slabs.slab_rows(0, 1, 6) == (0, 64)
    assert [slabs.slab_rows(r, 4, 6) for r in range(4)] == [(0, 16), (16, 16), (32, 16), (48, 16)]
    with pytest.raises(ValueError):
        slabs.slab_rows(0, 3, 6)
    with pytest.raises(ValueError):
        slabs.slab_rows(0, 64, 6)
    y = np.array([0, (16 << 26) - 1, 16 << 26, 0xFFFFFFFF], dtype=np.uint32)
    assert slabs.slab_of(y, 4, 6).tolist() == [0, 0, 1, 3]


def test_balance_rows_cuts_a_lopsided_scene_into_equal_shares():
    """psim_balance_rows (host code of libpsim_b200.so, no GPU): SURVEY.md section 8e's movable slab boundaries."""
    from particle_simulator_b200.stepper import PsimError, balance_rows

    w = workloads.clustered_mixed((8, 8), clusters=4, side=30, gas=2000, seed=5)
    p = w.frame.particles
    equal = np.bincount(slabs.slab_of(p["y"], 8, 8), minlength=8)
    bounds = balance_rows(w.frame, 8, 8)
    assert bounds[0] == 0 and bounds[-1] == 256 and (np.diff(bounds) >= 2).all()
    held = np.bincount(slabs.slab_of(p["y"], 8, 8, bounds), minlength=8)
    assert held.sum() == len(p)
    assert equal.max() / equal.mean() > 1.8                  # equal rows: badly balanced
    row_max = np.bincount((p["y"] >> np.uint32(24)).astype(np.int64)).max()
    assert held.max() - held.mean() <= row_max               # balanced to within one cell row of particles
    assert held.max() / held.mean() < 1.25, held
    parts = slabs.split_by_slab(p, 8, 8, bounds)
    assert [len(q) for q in parts] == held.tolist()
    # null records do not count; an empty scene still gives every slab its two rows
    q = p.copy()
    q["ty"][q["y"] < (1 << 31)] = -1
    half = FrameBuffer(len(q), w.frame.metadata)
    half.set_particles(q)
    b2 = balance_rows(half, 8, 4)
    assert b2[1] >= 128 - 2 and (np.diff(b2) >= 2).all()
    assert balance_rows(FrameBuffer(1), 6, 32) == list(range(0, 65, 2))
    assert balance_rows(w.frame, 8, 1) == [0, 256]
    with pytest.raises(PsimError, match="every slab needs 2"):
        balance_rows(w.frame, 6, 33)


def test_split_by_slab_is_a_partition_in_input_order():
    fb = FrameBuffer(4000)
    io.scene_hex_square(fb, 50, 80, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=11)
    p = fb.particles.copy()
    p["ty"][::97] = -1  # null records are nobody's
    parts = slabs.split_by_slab(p, 4, 6)
    assert sum(len(q) for q in parts) == int((p["ty"] >= 0).sum())
    for r, q in enumerate(parts):
        assert (slabs.slab_of(q["y"], 4, 6) == r).all()
    merged = np.concatenate(parts)
    order = np.argsort(slabs.slab_of(p[p["ty"] >= 0]["y"], 4, 6), kind="stable")
    assert merged.tobytes() == p[p["ty"] >= 0][order].tobytes()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_slab_crystal_pieces_make_one_crystal(world):
    """Each rank generates only its rows (+1 either side); what the ranks KEEP partitions the global crystal."""
    small = dict(per_slab=20_000, rows_per_slab_log2=6, grid_x_log2=7)
    geo = workloads.slab_crystal_geometry(world, **small)
    ly = geo["grid_log2"][1]
    whole = FrameBuffer(geo["nx"] * geo["ny"])
    whole.metadata["box_width"], whole.metadata["box_height"] = geo["box"]
    io.scene_hex_rows(whole, geo["nx"], geo["ny"], (0, geo["ny"]), (geo["box"][0] / 2, geo["box"][1] / 2), 1.0, 1.0,
                      10.0, 0, 3)
    want = slabs.split_by_slab(whole.particles, world, ly)
    for rank in range(world):
        wl = workloads.slab_crystal(rank, world, **small)
        assert wl.grid_log2 == geo["grid_log2"]
        kept = slabs.split_by_slab(wl.frame.particles, world, ly)[rank]
        assert kept.tobytes() == want[rank].tobytes()
        assert abs(len(kept) - 20_000) < 0.05 * 20_000  # equal work per slab


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = slabs.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else None, 128)
        assert uid == bytes(range(128))
        assert slabs.reduce_scalar(dist, 10.0 + rank, "max") == 10.0 + world - 1
        assert slabs.reduce_scalar(dist, 1.0 + rank, "sum") == sum(1.0 + r for r in range(world))
        # every rank builds its slab of the same crystal; the kept counts add up to the crystal
        small = dict(per_slab=5_000, rows_per_slab_log2=5, grid_x_log2=6)
        geo = workloads.slab_crystal_geometry(world, **small)
        wl = workloads.slab_crystal(rank, world, **small)
        kept = slabs.split_by_slab(wl.frame.particles, world, geo["grid_log2"][1])[rank]
        total = slabs.reduce_scalar(dist, float(len(kept)), "sum")
        assert total == geo["nx"] * geo["ny"]
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


class _HostSlab:
    """Stands in for a Stepper in the gloo test of slabs.rebalance_across_ranks: holds records, no GPU."""

    def __init__(self, grid_log2, meta, particles=None):
        self.grid_log2, self._meta = grid_log2, meta
        self.particles = particles
        self.closed = False

    def get_metadata(self):
        return self._meta

    def sync(self):
        pass

    def snapshot_async(self):
        pass

    def download(self):
        fb = FrameBuffer(max(len(self.particles), 1), self._meta)
        fb.set_particles(self.particles)
        return fb

    def upload(self, frame):
        self.particles = frame.particles.copy()

    def close(self):
        self.closed = True


def _rebalance_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = workloads.clustered_mixed((8, 8), clusters=4, side=30, gas=2000, seed=5)
        p = w.frame.particles
        p = p[np.argsort(p["y"] >> np.uint32(24), kind="stable")]  # rows ascending, like a cell-sorted state
        mine = slabs.split_by_slab(p, world, 8)[rank]              # equal rows: lopsided
        old = _HostSlab((8, 8), w.frame.metadata, mine)
        made = {}

        def make(bounds):
            made["bounds"] = list(bounds)
            return _HostSlab((8, 8), w.frame.metadata)

        new, bounds = slabs.rebalance_across_ranks(dist, old, make)
        assert old.closed and bounds == made["bounds"]
        from particle_simulator_b200.stepper import balance_rows

        assert bounds == balance_rows(w.frame, 8, world)  # the cut from the all-reduced histogram = the cut of the scene
        want = slabs.split_by_slab(p, world, 8, bounds)[rank]
        assert new.particles.tobytes() == want.tobytes()  # every record reached the owner of its row, order kept
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(len(want)))
    finally:
        dist.destroy_process_group()


def test_rebalance_across_ranks_routes_records_to_their_new_owners(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_rebalance_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    held = [int(open(os.path.join(tmp_path, f"ok{r}")).read()) for r in range(2)]
    assert max(held) / (sum(held) / 2) < 1.1, held
