"""step_kernel_c (particle_simulator_b200/csrc/step_float.cuh): the step kernel fine grids (>= 1024 cells per
axis) run, which stages stencil neighbours as exact fp32 offsets from zone / tile origins.

Checked (1) against the oracle restatement of the reference's bucket_step_kernel (kernel_bucket.cuh:40-94) on a
1024 x 1024 grid, and (2) against this library's integer-separation kernel (the one coarse grids run, forced
with PSIM_FORCE_INT_PATH=1) on scenes built to hit the fp32 path's corner cases: zone seams, the first and
last cell columns and rows, sparse tiles that fall back to global memory, crowded
cells, drifted (stale-membership) particles, walls, a 2:1 box on an 8192 x 4096 grid.
The two kernels compute the same separations exactly and differ only in rounding of the force law, so they
must agree to the L2 tolerance of test_gpu_parity (1e-5 of the largest pair force).
"""
import os

import numpy as np
import pytest

from conftest import frame_from
from oracle.oracle import PortOracle
from particle_simulator_b200 import FrameBuffer, default_metadata, io
from particle_simulator_b200.frame import PARTICLE_DTYPE
from test_gpu_parity import assert_state_close

pytestmark = pytest.mark.gpu

CELL = 50e-9 / 64  # the reference's cell width (kernel.cuh:15-18, particle.rs:141-142)


def run_steps(fb: FrameBuffer, grid, steps: int, force_int: bool, rebin_after: int = 0):
    """State after ingest and after `steps` steps (optionally a re-bin and one more step) on the chosen kernel."""
    from particle_simulator_b200.stepper import Stepper

    os.environ["PSIM_FORCE_INT_PATH"] = "1" if force_int else "0"
    try:
        with Stepper(grid, max(fb.count, 1)) as st:
            st.upload(fb)
            before = st.download().particles.copy()
            st.step_async(steps)
            if rebin_after:
                st.rebin_async()
                st.step_async(rebin_after)
            st.snapshot_async()
            return before, st.download().particles.copy()
    finally:
        os.environ.pop("PSIM_FORCE_INT_PATH", None)


def tile_stats(fb: FrameBuffer, grid, force_int: bool = False) -> dict:
    from particle_simulator_b200.stepper import Stepper

    os.environ["PSIM_FORCE_INT_PATH"] = "1" if force_int else "0"
    try:
        with Stepper(grid, max(fb.count, 1)) as st:
            st.upload(fb)
            return st.tile_stats()
    finally:
        os.environ.pop("PSIM_FORCE_INT_PATH", None)


def boxed(n: int, grid, meta=None) -> FrameBuffer:
    fb = FrameBuffer(n, meta)
    fb.metadata["box_width"] = CELL * (1 << grid[0])
    fb.metadata["box_height"] = CELL * (1 << grid[1])
    return fb


def assert_paths_agree(fb, grid, steps=1, what="", rebin_after=0):
    b_f, a_f = run_steps(fb, grid, steps, force_int=False, rebin_after=rebin_after)
    b_i, a_i = run_steps(fb, grid, steps, force_int=True, rebin_after=rebin_after)
    assert b_f.tobytes() == b_i.tobytes()
    if steps == 1 and not rebin_after:
        assert_state_close(a_f, a_i, b_f, fb.metadata, what)
    else:  # errors compound over steps: compare position by position, loosely
        assert np.array_equal(a_f["ty"], a_i["ty"])
        dx = np.abs((a_f["x"].astype(np.int64) - a_i["x"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        dy = np.abs((a_f["y"].astype(np.int64) - a_i["y"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        assert max(dx.max(), dy.max()) <= 64 * steps, (what, dx.max(), dy.max())
        # a hot liquid amplifies rounding differences step by step: velocities agree to ~1e-4 of the thermal speed
        assert np.allclose(a_f["vx"], a_i["vx"], rtol=1e-3, atol=0.05) and np.allclose(a_f["vy"], a_i["vy"], rtol=1e-3, atol=0.05)
    return a_f


def test_liquid_patch_on_1024_grid_vs_oracle():
    """zl = 1 (zones of two columns): 160k particles at liquid density, one step, against the oracle."""
    grid = (10, 10)
    fb = boxed(400 * 400, grid)
    w = float(fb.metadata["box_width"])
    io.scene_hex_square(fb, 400, 400, (0.37 * w, 0.61 * w), 1.06, 100.0, 200.0, 0, seed=21)
    port = PortOracle(10, 10, 16)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    want = port.step(slots, fb.metadata, threads=8)
    want = want[want["ty"] >= 0]
    before, got = run_steps(fb, grid, 1, force_int=False)
    assert before.tobytes() == slots[slots["ty"] >= 0].tobytes()
    assert_state_close(got, want, before, fb.metadata, "liquid on 1024^2")
    st = tile_stats(fb, grid)  # it really was step_kernel_c, every tile staged in shared memory
    assert st["float_path"] == 1 and st["tiles"] > 300 and st["tiles_staged"] >= 0.99 * st["tiles"]
    assert tile_stats(fb, grid, force_int=True)["float_path"] == 0


@pytest.mark.parametrize("grid", [(10, 10), (11, 11), (12, 10), (13, 12)])
def test_lattice_across_zone_seams_and_grid_edges(grid):
    """A crystal that starts in cell column 0 / row 0 region and spans many zones, plus one in the far corner."""
    n1, n2 = 300, 120
    fb = boxed(n1 * n1 + n2 * n2, grid)
    w, h = float(fb.metadata["box_width"]), float(fb.metadata["box_height"])
    a = FrameBuffer(n1 * n1, fb.metadata)
    io.scene_hex_square(a, n1, n1, (n1 * 2.05e-10 + 1.2e-9, n1 * 1.8e-10 + 1.2e-9), 1.0, 1.0, 30.0, 0, seed=22)
    b = FrameBuffer(n2 * n2, fb.metadata)
    io.scene_hex_square(b, n2, n2, (w - n2 * 2.05e-10 - 1.2e-9, h - n2 * 1.8e-10 - 1.2e-9), 1.03, 1.0, 30.0, 1, seed=23)
    fb.set_particles(np.concatenate([a.particles, b.particles]))
    assert_paths_agree(fb, grid, what=f"seams {grid}")


def test_gas_tiles_straddle_rows_and_sparse_tiles_fall_back():
    """2.4 particles per cell in one corner (tiles span ~50 cells, some straddle two rows) and a thin gas
    elsewhere (tiles span many rows: global-memory fallback inside the same kernel)."""
    grid = (10, 10)
    rng = np.random.default_rng(31)
    w = CELL * 1024
    # dense corner: a jittered square lattice of 0.52 nm (2.25 per cell, nearest pairs >= 1.1 sigma)
    gx, gy = np.meshgrid(np.arange(240), np.arange(240))
    dense = np.stack([gx.ravel(), gy.ravel()], axis=1) * 5.2e-10 + 2 * CELL + rng.uniform(-6e-11, 6e-11, (240 * 240, 2))
    # thin gas elsewhere: a jittered 6 nm lattice, 0.017 per cell
    tx, ty = np.meshgrid(np.arange(130), np.arange(130))
    thin = np.stack([tx.ravel(), ty.ravel()], axis=1) * 6e-9 + 4e-9 + rng.uniform(-2e-9, 2e-9, (130 * 130, 2))
    thin = thin[(thin[:, 0] > 130e-9) | (thin[:, 1] > 130e-9)]
    xy = np.concatenate([dense, thin])
    assert xy.min() > 1e-9 and xy.max() < w - 1e-9
    p = np.zeros(len(xy), dtype=PARTICLE_DTYPE)
    p["x"] = np.round(xy[:, 0] / w * 2**32).astype(np.uint64).astype(np.uint32)
    p["y"] = np.round(xy[:, 1] / w * 2**32).astype(np.uint64).astype(np.uint32)
    v = rng.normal(0, 200.0, (len(xy), 2))
    p["vx"], p["vy"] = v[:, 0], v[:, 1]
    fb = boxed(len(p), grid)
    fb.set_particles(p)
    fb.metadata["step_dt"] = 10e-15
    port = PortOracle(10, 10, 32)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    _, _, max_pair = port.forces(slots, fb.metadata)
    want = port.step(slots, fb.metadata, threads=8)
    want = want[want["ty"] >= 0]
    before, got = run_steps(fb, grid, 1, force_int=False)
    assert_state_close(got, want, before, fb.metadata, "gas on 1024^2", max_pair[slots["ty"] >= 0])
    assert_paths_agree(fb, grid, what="gas, both kernels")
    st = tile_stats(fb, grid)  # the dense corner is staged, the thin gas is not
    assert st["float_path"] == 1 and 50 < st["tiles_staged"] < st["tiles"]


def test_stale_membership_and_rebin_on_fine_grid():
    """17 steps of a hot liquid without re-binning (particles drift out of their membership cells), a re-bin,
    and more steps: the fp32 path must track the integer kernel the whole way."""
    grid = (11, 11)
    fb = boxed(350 * 350, grid)
    w = float(fb.metadata["box_width"])
    io.scene_hex_square(fb, 350, 350, (0.5 * w + 3.1e-10, 0.25 * w), 1.05, 150.0, 250.0, 0, seed=24)
    fb.metadata["step_dt"] = 20e-15
    assert_paths_agree(fb, grid, steps=17, what="stale", rebin_after=5)


def test_crowded_cells():
    """A compressed blob, ~19 particles per cell (the reference's slot array holds 16): tiles get narrower
    instead of overflowing the band buffer."""
    grid = (10, 10)
    fb = boxed(90 * 90, grid)
    w = float(fb.metadata["box_width"])
    fb.metadata["step_dt"] = 1e-15
    io.scene_square(fb, 90, 90, (0.5 * w, 0.5 * w), 0.45, 0.0, 10.0, 0, seed=25)
    b_f, a_f = run_steps(fb, grid, 1, force_int=False)
    b_i, a_i = run_steps(fb, grid, 1, force_int=True)
    port = PortOracle(10, 10, 64)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    _, _, max_pair = port.forces(slots, fb.metadata)
    assert_state_close(a_f, a_i, b_f, fb.metadata, "crowded", max_pair[slots["ty"] >= 0])
    st = tile_stats(fb, grid)
    assert st["float_path"] == 1 and st["tiles_staged"] == st["tiles"]


def test_other_exponents_on_fine_grid():
    """Every compile-time power variant of step_kernel_c (kn = 5..10, with and without the cubic) vs the integer kernel;
    exponents it has no variant for run step_kernel on fine grids too."""
    grid = (10, 10)
    for n_exp in (8.1, 10.0, 12.085, 14.08, 16.0, 18.2, 14.3, 24.0):
        fb = boxed(150 * 150, grid)
        w = float(fb.metadata["box_width"])
        fb.metadata["particles"][0] = (3.609e-10, 1.46e-21, n_exp, 6.0)
        io.scene_hex_square(fb, 150, 150, (0.3 * w, 0.7 * w), 1.04, 50.0, 150.0, 0, seed=26)
        assert_paths_agree(fb, grid, what=f"n={n_exp}")


def test_walls_and_cursor_on_fine_grid():
    grid = (10, 10)
    meta = default_metadata()
    meta["cursor_pos"] = (0.02, 0.03)
    meta["cursor_size"] = 0.05
    fb = boxed(100 * 100, grid, meta)
    io.scene_hex_square(fb, 100, 100, (100 * 2.05e-10 + 6e-10, 100 * 1.8e-10 + 6e-10), 1.0, 1.0, 30.0, 0, seed=27)
    assert_paths_agree(fb, grid, what="walls")


def test_metadata_updates_between_steps_rebuild_the_neighbour_records():
    """The records carry the scale of sigma and exist only while the physics has an fp32 variant: a header-only
    metadata update (Kernel::write_metadata, kernel.cuh:96-101) changes sigma, then switches to exponents that run
    the integer kernel, then back. Same sequence on the integer kernel throughout; states must track each other."""
    from particle_simulator_b200.stepper import Stepper

    grid = (10, 10)
    fb = boxed(160 * 160, grid)
    w = float(fb.metadata["box_width"])
    io.scene_hex_square(fb, 160, 160, (0.4 * w, 0.55 * w), 1.04, 20.0, 60.0, 0, seed=28)
    metas = [fb.metadata.copy() for _ in range(4)]
    metas[1]["particles"][0]["sigma"] = np.float32(3.7e-10)        # another scale: (kx / sigma) changes
    metas[2]["particles"][0]["n"] = np.float32(24.0)               # no fp32 variant: step_kernel takes over
    metas[3]["particles"][0]["n"] = np.float32(12.0)               # back, with other powers

    def run(force_int: bool):
        os.environ["PSIM_FORCE_INT_PATH"] = "1" if force_int else "0"
        try:
            out, paths = [], []
            with Stepper(grid, fb.count) as st:
                st.upload(fb)
                for m in metas:
                    st.set_metadata(m)
                    st.step_async(3)
                    st.snapshot_async()
                    out.append(st.download().particles.copy())
                    paths.append(st.tile_stats()["float_path"])
            return out, paths
        finally:
            os.environ.pop("PSIM_FORCE_INT_PATH", None)

    got, paths = run(False)
    want, paths_int = run(True)
    assert paths == [1, 1, 0, 1] and paths_int == [0, 0, 0, 0]
    for k, (g, wnt) in enumerate(zip(got, want)):
        assert np.array_equal(g["ty"], wnt["ty"])
        dx = np.abs((g["x"].astype(np.int64) - wnt["x"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        dy = np.abs((g["y"].astype(np.int64) - wnt["y"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        assert max(dx.max(), dy.max()) <= 64 * 3 * (k + 1), (k, dx.max(), dy.max())
        assert np.allclose(g["vx"], wnt["vx"], rtol=1e-3, atol=0.05) and np.allclose(g["vy"], wnt["vy"], rtol=1e-3, atol=0.05)


def test_frames_on_fine_grids_do_not_wait_for_the_host_and_replay_as_graphs():
    """A single slab's re-bin leaves the tile count in device memory: psim_run_frame_async returns without waiting for
    the frame (the reference's compute_frame wanted exactly that, cuda_simulator.cu:7-26), frames can be captured as
    CUDA graphs on fine grids too, and a launch smaller than the tile count (PSIM_TILE_LAUNCH_CAP: the count outgrew
    the host's margin) still steps every tile. All bit-identical."""
    import time

    from particle_simulator_b200.stepper import Stepper

    grid = (10, 10)
    fb = boxed(500 * 500, grid)
    w = float(fb.metadata["box_width"])
    io.scene_hex_square(fb, 500, 500, (0.45 * w, 0.52 * w), 1.05, 150.0, 250.0, 0, seed=29)
    fb.metadata["step_dt"] = 20e-15
    fb.metadata["steps_per_frame"] = 100

    def run(use_graph: bool, launch_cap: int | None = None):
        if launch_cap:
            os.environ["PSIM_TILE_LAUNCH_CAP"] = str(launch_cap)
        try:
            with Stepper(grid, fb.count, use_graph=use_graph) as st:
                st.upload(fb)
                assert st.tile_stats()["float_path"] == 1
                frames, enqueue, total = [], [], []
                for _ in range(4):
                    st.sync()
                    t0 = time.perf_counter()
                    st.run_frame_async()
                    t1 = time.perf_counter()
                    st.sync()
                    t2 = time.perf_counter()
                    enqueue.append(t1 - t0)
                    total.append(t2 - t0)
                    frames.append(st.download().particles.copy())
                assert (st.steps_executed, st.rebins_executed) == (404, 24)
                return frames, enqueue, total
        finally:
            os.environ.pop("PSIM_TILE_LAUNCH_CAP", None)

    plain, enq_plain, tot_plain = run(False)
    graph, enq_graph, tot_graph = run(True)
    capped, _, _ = run(False, launch_cap=7)
    for a, b, c in zip(plain, graph, capped):
        assert a.tobytes() == b.tobytes() == c.tobytes()
    print(f"frame {1e3 * min(tot_plain):.2f} ms; enqueue {1e3 * min(enq_plain):.2f} ms launch by launch, "
          f"{1e3 * min(enq_graph[1:]):.3f} ms as a graph")
    assert min(enq_graph[1:]) < 0.25 * min(tot_graph)  # the host is free while the frame runs
