"""Slab decomposition over cell rows on ONE GPU: the in-process slab group
(psim_group_*, include/psim_b200.h) runs the same slabs, per-step halo exchange and re-bin migration
as the one-process-per-GPU NCCL mode, with device-to-device copies as the transport.

The reference is single-device, so parity here is with our own single-slab run, and it is BIT-EXACT:
every particle accumulates its stencil neighbours in the same order whatever slab it lives in, and
migrants are merged in the order they have in the global cell-sorted array.
"""
import numpy as np
import pytest

from conftest import frame_from
from particle_simulator_b200 import FrameBuffer, io

pytestmark = pytest.mark.gpu


def run_both(fb: FrameBuffer, grid, slabs: int, frames: int, per_slab_capacity: int | None = None, bounds=None):
    """Run `frames` frames single-slab and as a slab group; yield (single, group, the group) per frame."""
    from particle_simulator_b200.stepper import SlabGroup, Stepper

    n = fb.count
    cap = per_slab_capacity or n
    with Stepper(grid, n) as st, SlabGroup(grid, slabs, cap, ingest_capacity=n, bounds=bounds) as gr:
        st.upload(fb)
        gr.upload(fb)
        assert gr.particle_count == st.particle_count
        assert st.download().particles.tobytes() == gr.download().particles.tobytes()  # the ingested scene
        for _ in range(frames):
            st.run_frame_async()
            gr.run_frame_async()
            st.sync()
            gr.sync()
            yield st.download().particles.copy(), gr.download().particles.copy(), gr
        assert [s.steps_executed for s in gr.slabs] == [st.steps_executed] * slabs
        assert [s.rebins_executed for s in gr.slabs] == [st.rebins_executed] * slabs


@pytest.mark.parametrize("slabs", [2, 4, 8])
@pytest.mark.parametrize("scene", ["hex2500", "gas10k", "wall_cursor"])
def test_group_is_bit_identical_to_single_slab(golden, scene, slabs):
    g = golden(scene)
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 35  # 36 steps, 2 re-bins per frame on the reference schedule
    fb = frame_from(g["input"], meta)
    stages = 0
    for single, group, _ in run_both(fb, (6, 6), slabs, frames=3):
        assert len(single) == len(group) == fb.count
        assert single.tobytes() == group.tobytes()
        stages += 1
    assert stages == 3


@pytest.mark.parametrize("scene,bounds", [("hex2500", [0, 30, 32, 34, 64]),          # two 2-row slabs in the crystal
                                          ("gas10k", [0, 2, 11, 40, 62, 64]),       # 5 slabs (not a power of two)
                                          ("wall_cursor", [0, 21, 64])])
def test_slabs_of_unequal_heights_are_bit_identical_too(golden, scene, bounds):
    """PsimConfig.slab_bounds: every slab owns the rows it is told to."""
    g = golden(scene)
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 35
    fb = frame_from(g["input"], meta)
    for single, group, gr in run_both(fb, (6, 6), len(bounds) - 1, frames=3, bounds=bounds):
        assert single.tobytes() == group.tobytes()
        info = [s.slab_info() for s in gr.slabs]
        assert [i["first_row"] for i in info] == bounds[:-1] and [i["rows"] for i in info] == np.diff(bounds).tolist()


def test_balanced_boundaries_and_rebalancing_a_running_group():
    """the clustered scene (BASELINE.json configs[4]) cut by psim_balance_rows instead of into
    equal numbers of rows, on a fine grid (step_kernel_c with the halo pushed by the step kernel). Results stay
    bit-identical to the single slab; the boundaries are moved again while the scene runs."""
    from particle_simulator_b200 import workloads
    from particle_simulator_b200.stepper import SlabGroup, Stepper, balance_rows

    w = workloads.clustered_mixed((10, 10), clusters=6, side=150, gas=20000, seed=5)
    w.frame.metadata["steps_per_frame"] = 50
    n = w.particles
    bounds = balance_rows(w.frame, 10, 8)
    with Stepper(w.grid_log2, n) as st, SlabGroup(w.grid_log2, 8, n, ingest_capacity=n, bounds=bounds) as gr:
        st.upload(w.frame)
        gr.upload(w.frame)
        held = np.array([s.particle_count for s in gr.slabs])
        assert held.sum() == n and held.max() / held.mean() < 1.1, held   # equal rows: 2.4 (test_gpu_configs.py)
        assert gr.slabs[3].tile_stats()["float_path"] == 1
        for frame in range(3):
            st.run_frame_async()
            gr.run_frame_async()
            st.sync()
            gr.sync()
            assert st.download().particles.tobytes() == gr.download().particles.tobytes()
            if frame == 1:
                # move the boundaries to where the particles are now; the single slab restarts its schedule the same
                # way by taking its own state back in
                new = gr.rebalance()
                assert (np.diff(new) >= 2).all() and new[0] == 0 and new[-1] == 1024
                st.snapshot_async()
                st.upload(st.download())
                held = np.array([s.particle_count for s in gr.slabs])
                assert held.sum() == n and held.max() / held.mean() < 1.1, held
                assert [s.slab_info()["first_row"] for s in gr.slabs] == new[:-1]


def test_metadata_change_between_frames_on_pushed_halo_fine_grid():
    """A header-only metadata update (Kernel::write_metadata, kernel.cuh:96-101) that changes the scale of the
    neighbour records (sigma, then the box) while the ghost rows are being filled by the neighbours' pushes: the
    slabs rebuild their own rows and re-deliver the ghost rows (team_refresh_stale_records, stepper.cu). Bit-identical
    to the single slab, which rebuilds everything locally."""
    from particle_simulator_b200 import workloads
    from particle_simulator_b200.stepper import SlabGroup, Stepper

    w = workloads.clustered_mixed((10, 10), clusters=4, side=120, gas=8000, seed=7)
    w.frame.metadata["steps_per_frame"] = 20
    n = w.particles
    metas = [w.frame.metadata.copy() for _ in range(3)]
    metas[1]["particles"][0]["sigma"] = np.float32(3.7e-10)
    metas[2]["box_width"] = np.float32(float(metas[2]["box_width"]) * 1.01)
    metas[2]["box_height"] = np.float32(float(metas[2]["box_height"]) * 1.01)
    with Stepper(w.grid_log2, n) as st, SlabGroup(w.grid_log2, 4, n, ingest_capacity=n) as gr:
        st.upload(w.frame)
        gr.upload(w.frame)
        assert gr.slabs[1].tile_stats()["float_path"] == 1 and gr.slabs[1].halo_mode == 2
        for m in metas:
            st.set_metadata(m)
            gr.set_metadata(m)
            for _ in range(2):
                st.run_frame_async()
                gr.run_frame_async()
                st.sync()
                gr.sync()
                assert st.download().particles.tobytes() == gr.download().particles.tobytes()


def test_ingest_and_fine_grained_calls_match(golden):
    from particle_simulator_b200.stepper import SlabGroup, Stepper

    g = golden("liquid4k")
    fb = frame_from(g["input"], g["meta"][0])
    with Stepper((6, 6), fb.count) as st, SlabGroup((6, 6), 4, fb.count, ingest_capacity=fb.count) as gr:
        st.upload(fb)
        gr.upload(fb)
        # the ingested scene: golden binning of the reference (kernel.cuh:210-239), slab by slab
        assert gr.download().particles.tobytes() == g["binned"].tobytes()
        rows = [s.slab_info() for s in gr.slabs]
        assert [r["first_row"] for r in rows] == [0, 16, 32, 48] and all(r["rows"] == 16 for r in rows)
        assert [r["local_rows"] for r in rows] == [17, 18, 18, 17]
        assert sum(r["particles"] for r in rows) == fb.count
        # a slab's ghost rows are its neighbours' boundary rows
        cs = [s.cell_start() for s in gr.slabs]
        for r in range(3):
            top_row_of_r = np.diff(cs[r])[-2 * 64:-64]
            ghost_below_next = np.diff(cs[r + 1])[:64]
            assert np.array_equal(top_row_of_r, ghost_below_next)
        for k in range(1, 40):
            st.step_async(1)
            gr.step_async(1)
            if k % 9 == 0:
                st.rebin_async()
                gr.rebin_async()
            st.snapshot_async()
            gr.snapshot_async()
            assert st.download().particles.tobytes() == gr.download().particles.tobytes(), f"step {k}"


def test_particles_migrate_between_slabs(golden):
    """The hot gas crosses slab boundaries: ownership moves, nothing is lost or duplicated."""
    g = golden("gas10k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 100
    fb = frame_from(g["input"], meta)
    counts = []
    for single, group, gr in run_both(fb, (6, 6), 4, frames=4):
        assert single.tobytes() == group.tobytes()
        counts.append([s.particle_count for s in gr.slabs])
        assert sum(counts[-1]) == fb.count
    assert len({tuple(c) for c in counts}) > 1, counts  # ownership really changed


@pytest.mark.parametrize("drift", [2500.0, -2500.0])
def test_thousands_of_migrants_per_rebin(drift):
    """The whole 1M liquid drifts up (down) at 2.5 km/s: about a thousand migrants per slab boundary and re-bin (half a
    cell row in 17 steps), ranked by the block counts + ballots of the migrant kernels. Still bit-identical to the single
    slab, nobody lost."""
    from particle_simulator_b200.workloads import config_1m_liquid

    w = config_1m_liquid()
    w.frame.metadata["steps_per_frame"] = 52  # 52 steps, 3 re-bins
    w.frame.particles["vy"] += np.float32(drift)
    sent = 0
    for single, group, gr in run_both(w.frame, w.grid_log2, 8, frames=2, per_slab_capacity=w.particles // 2):
        assert single.tobytes() == group.tobytes()
        assert sum(s.particle_count for s in gr.slabs) == w.particles
        sent = sum(s.migrants_sent for s in gr.slabs)
    assert sent > 10000, sent


def test_liquid_1m_in_8_slabs():
    """configs[1]-sized: 1M particles melting at 150-250 m/s, 1024x1024 cells, 8 slabs of 128 rows."""
    from particle_simulator_b200.workloads import config_1m_liquid

    w = config_1m_liquid()
    w.frame.metadata["steps_per_frame"] = 52  # 52 steps, 3 re-bins
    n = 0
    for single, group, gr in run_both(w.frame, w.grid_log2, 8, frames=2, per_slab_capacity=w.particles // 2):
        assert single.tobytes() == group.tobytes()
        n += 1
    assert n == 2


def test_group_with_copies_after_every_step_matches_too(golden, monkeypatch):
    """PSIM_HALO=nccl switches the in-process group from pushes by the step kernel to a device-to-device copy of
    the boundary rows after every step (the shape of the send/recv exchange)."""
    from particle_simulator_b200.stepper import SlabGroup

    g = golden("gas10k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 35
    fb = frame_from(g["input"], meta)
    with SlabGroup((6, 6), 4, fb.count, ingest_capacity=fb.count) as gr:
        assert [s.halo_mode for s in gr.slabs] == [2] * 4
    monkeypatch.setenv("PSIM_HALO", "nccl")
    n = 0
    for single, group, gr in run_both(fb, (6, 6), 4, frames=2):
        assert [s.halo_mode for s in gr.slabs] == [1] * 4
        assert single.tobytes() == group.tobytes()
        n += 1
    assert n == 2


def test_fine_grid_slabs_run_the_couples_kernel():
    """1024 x 1024 cells: the slabs run step_kernel_c; its tiles of the first / last owned row push the halo."""
    from particle_simulator_b200.workloads import lattice

    w = lattice(300, 300, (10, 10), 1.04, 100.0, 200.0, seed=5)
    w.frame.metadata["steps_per_frame"] = 35
    n = 0
    for single, group, gr in run_both(w.frame, w.grid_log2, 4, frames=2):
        assert single.tobytes() == group.tobytes()
        assert all(s.tile_stats()["float_path"] == 1 for s in gr.slabs)
        n += 1
    assert n == 2


def test_native_schedule_matches_too(golden):
    from particle_simulator_b200.stepper import SCHEDULE_NATIVE, SlabGroup, Stepper

    g = golden("gas10k")
    meta = g["meta"][0].copy()
    meta["steps_per_frame"] = 25
    fb = frame_from(g["input"], meta)
    with Stepper((6, 6), fb.count, schedule=SCHEDULE_NATIVE, rebin_every=7) as st, \
            SlabGroup((6, 6), 2, fb.count, ingest_capacity=fb.count, schedule=SCHEDULE_NATIVE, rebin_every=7) as gr:
        st.upload(fb)
        gr.upload(fb)
        for _ in range(3):
            st.run_frame_async()
            gr.run_frame_async()
            st.sync()
            gr.sync()
            assert st.download().particles.tobytes() == gr.download().particles.tobytes()
        assert st.steps_executed == 75 and gr.slabs[0].steps_executed == 75
        assert st.rebins_executed == gr.slabs[1].rebins_executed == 10


def test_particle_that_skips_a_slab_is_reported_not_lost():
    from particle_simulator_b200.stepper import PsimError, SlabGroup

    fb = FrameBuffer(2501)
    io.scene_hex_square(fb, 50, 50, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=7)
    p = fb.particles[:1].copy()
    p["x"], p["y"], p["vx"], p["vy"] = 1 << 28, 1 << 28, 0.0, 40000.0  # 34 nm in 17 steps = 43 cell rows
    fb.set_particles(np.concatenate([fb.particles, p]))
    fb.metadata["steps_per_frame"] = 36  # the second re-bin comes 17 steps after the first
    with SlabGroup((6, 6), 8, 2501, ingest_capacity=2501) as gr:
        gr.upload(fb)
        with pytest.raises(PsimError, match="moved past the adjacent slab"):
            gr.run_frame_async()
            gr.sync()


def test_bad_slab_configurations_are_rejected():
    from particle_simulator_b200.stepper import PsimError, Stepper

    with pytest.raises(PsimError, match="do not split"):
        Stepper((6, 6), 100, slab_rank=0, slab_count=3)
    with pytest.raises(PsimError, match="slab_rank"):
        Stepper((6, 6), 100, slab_rank=2, slab_count=2)
    for bad in ([0, 1, 64], [0, 63, 64], [0, 40, 30, 64], [2, 30, 64], [0, 30, 60]):
        with pytest.raises(PsimError, match="slab_bounds"):
            for r in range(len(bad) - 1):
                Stepper((6, 6), 100, slab_rank=r, slab_count=len(bad) - 1, bounds=bad).close()
    with Stepper((6, 6), 100, slab_rank=1, slab_count=2) as st:
        fb = FrameBuffer(4)
        io.scene_hex_square(fb, 2, 2, (25e-9, 40e-9), 1.0, 5.0, 5.0, 0, seed=7)
        with pytest.raises(PsimError, match="psim_comm_init"):
            st.upload(fb)  # a lone slab of two must have joined its communicator first
