"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle
would need minutes per step there): conservation of the particle set, sortedness and consistency
of the binning, idempotence of the re-bin, momentum conservation of the pair forces, and agreement
of a 1M-particle window of the 10M run with the oracle on that window."""
import numpy as np
import pytest

from particle_simulator_b200 import FrameBuffer, io

pytestmark = pytest.mark.gpu


def cells_of(p, lx, ly):
    return (p["x"] >> np.uint32(32 - lx)).astype(np.int64) + ((p["y"] >> np.uint32(32 - ly)).astype(np.int64) << lx)


@pytest.mark.parametrize("n_side,grid,box", [(1000, (10, 10), 0.8e-6),      # configs[1]: 1M liquid
                                             (3162, (11, 11), 1.6e-6)])     # configs[2]: 10M lattice
def test_full_size_frame_properties(n_side, grid, box):
    from particle_simulator_b200.stepper import Stepper

    n = n_side * n_side
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = box
    fb.metadata["box_height"] = box
    fb.metadata["steps_per_frame"] = 18  # 18 steps, 1 re-bin
    io.scene_hex_square(fb, n_side, n_side, (box / 2, box / 2), 1.05 if n_side == 1000 else 1.0, 1.0, 10.0, 0, seed=3)
    lx, ly = grid
    with Stepper(grid, n) as st:
        st.upload(fb)
        assert st.particle_count == n
        cs = st.cell_start()
        binned = st.download().particles.copy()
        # ingest: a permutation of the input, sorted by cell, stable, counts consistent
        c = cells_of(binned, lx, ly)
        assert (np.diff(c) >= 0).all()
        assert np.array_equal(np.diff(cs), np.bincount(c, minlength=1 << (lx + ly)))
        want = fb.particles[np.argsort(cells_of(fb.particles, lx, ly), kind="stable")]
        assert binned.tobytes() == want.tobytes()
        # one frame
        st.run_frame_async()
        st.sync()
        assert (st.steps_executed, st.rebins_executed) == (18, 1)
        out = st.download().particles.copy()
        assert len(out) == n and (out["ty"] == 0).all()
        assert np.isfinite(out["vx"]).all() and np.isfinite(out["vy"]).all()
        # pair forces conserve momentum; the lattice is far from the walls
        p0 = np.array([binned["vx"].astype(np.float64).sum(), binned["vy"].astype(np.float64).sum()])
        p1 = np.array([out["vx"].astype(np.float64).sum(), out["vy"].astype(np.float64).sum()])
        vsum = np.abs(out["vx"].astype(np.float64)).sum()
        assert np.abs(p1 - p0).max() < 1e-5 * vsum
        # re-bin: sorted, idempotent, a permutation
        st.rebin_async()
        st.snapshot_async()
        a = st.download().particles.copy()
        assert (np.diff(cells_of(a, lx, ly)) >= 0).all()
        assert a.tobytes() == out[np.argsort(cells_of(out, lx, ly), kind="stable")].tobytes()
        st.rebin_async()
        st.snapshot_async()
        assert st.download().particles.tobytes() == a.tobytes()


def test_10m_window_single_step_vs_oracle():
    """One step of the 10M-particle lattice; a 256x256-cell window of it is re-run through the oracle
    (particles of the window plus a one-cell rim, shifted to the oracle's own origin)."""
    from oracle.oracle import PortOracle
    from particle_simulator_b200.stepper import Stepper
    from test_gpu_parity import assert_state_close

    n_side, box, lx = 3162, 1.6e-6, 11
    n = n_side * n_side
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = box
    fb.metadata["box_height"] = box
    io.scene_hex_square(fb, n_side, n_side, (box / 2, box / 2), 1.0, 1.0, 10.0, 0, seed=3)
    with Stepper((lx, lx), n) as st:
        st.upload(fb)
        before = st.download().particles.copy()
        st.step_async(1)
        st.snapshot_async()
        after = st.download().particles.copy()
    # window: cells [896, 1152) in both axes = a 256x256 block in the middle of the crystal.
    # Same cell width in a 256-cell oracle grid with box/8; fixed-point coordinates scale by 8.
    lo, hi = 896, 1152
    cx = (before["x"] >> np.uint32(32 - lx)).astype(np.int64)
    cy = (before["y"] >> np.uint32(32 - lx)).astype(np.int64)
    sel = (cx >= lo) & (cx < hi) & (cy >= lo) & (cy < hi)
    sub = before[sel].copy()
    origin = np.uint32(lo << (32 - lx))
    sub["x"] = (sub["x"] - origin) << np.uint32(3)
    sub["y"] = (sub["y"] - origin) << np.uint32(3)
    wfb = FrameBuffer(len(sub))
    wfb.metadata["box_width"] = box / 8
    wfb.metadata["box_height"] = box / 8
    wfb.set_particles(sub)
    port = PortOracle(8, 8, 16)
    slots, dropped = port.prepare(wfb)
    assert dropped == 0
    want = port.step(slots, wfb.metadata, threads=8)
    want = want[want["ty"] >= 0]
    got = after[sel].copy()
    got["x"] = (got["x"] - origin) << np.uint32(3)
    got["y"] = (got["y"] - origin) << np.uint32(3)
    # compare the interior only: the rim of the window misses its outside neighbours in the oracle run,
    # and feels the oracle box's walls
    wcx = (sub["x"] >> np.uint32(24)).astype(np.int64)
    wcy = (sub["y"] >> np.uint32(24)).astype(np.int64)
    inner = (wcx >= 8) & (wcx < 248) & (wcy >= 8) & (wcy < 248)
    assert inner.sum() > 200_000
    # the window is cell-sorted in the same relative order as the oracle's own binning of it
    assert np.array_equal(slots[slots["ty"] >= 0]["vx"], sub["vx"])
    # positions were scaled by 8 (exact), so LSB slack scales too: compare in window units / 8
    g, w, b = got[inner].copy(), want[inner].copy(), sub[inner].copy()
    for a in (g, w, b):
        a["x"] >>= np.uint32(3)
        a["y"] >>= np.uint32(3)
    assert_state_close(g, w, b, wfb.metadata, "10M window")


def test_100m_liquid_on_one_gpu():
    """BASELINE.json configs[3]'s size on ONE B200 (the 8-GPU decomposition of the same box is what bench.py --gpus 8
    runs per slab): 98 M particles at liquid spacing on 8192 x 4096 cells (box 6.4 x 3.2 um, a 2:1 box: the fp32
    path folds ky/kx = 1/2 into its scale). One frame of 18 steps and a re-bin; conservation and consistency only."""
    from particle_simulator_b200.stepper import Stepper

    nx, ny, grid = 14000, 7000, (13, 12)
    n = nx * ny
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = 6.4e-6
    fb.metadata["box_height"] = 3.2e-6
    fb.metadata["steps_per_frame"] = 18
    io.scene_hex_square(fb, nx, ny, (3.2e-6, 1.6e-6), 1.05, 1.0, 10.0, 0, seed=4)
    with Stepper(grid, n) as st:
        st.upload(fb)
        assert st.particle_count == n
        stats = st.tile_stats()
        assert stats["float_path"] == 1 and stats["tiles_staged"] >= 0.99 * stats["tiles"]
        cs = st.cell_start()
        assert cs[0] == 0 and cs[-1] == n and (np.diff(cs.astype(np.int64)) >= 0).all()
        p0 = np.array([fb.particles["vx"].sum(dtype=np.float64), fb.particles["vy"].sum(dtype=np.float64)])
        st.run_frame_async()
        st.sync()
        assert (st.steps_executed, st.rebins_executed) == (18, 1)
        out = st.download(fb)  # into the same host buffer
    assert out.count == n
    p = out.particles
    assert (p["ty"] == 0).all() and np.isfinite(p["vx"]).all() and np.isfinite(p["vy"]).all()
    c = cells_of(p, *grid)
    assert (np.diff(c) >= -(1 << 13)).all()  # still (nearly) cell-sorted: a snapshot is not freshly binned
    p1 = np.array([p["vx"].sum(dtype=np.float64), p["vy"].sum(dtype=np.float64)])
    assert np.abs(p1 - p0).max() < 1e-5 * np.abs(p["vx"]).sum(dtype=np.float64)  # pair forces conserve momentum
