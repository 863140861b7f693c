"""PsimConfig.species_physics: per-species Mie parameters, an EXTENSION -- the reference steps
every particle with metadata.particles[0] (kernel_bucket.cuh:52). The GPU kernel (step_kernel_species, csrc/step_int.cuh)
is checked against the same extension of the CPU restatement (oracle_step_species, oracle/psim_oracle.c), whose
species-0-only case is the pinned reference step."""
import numpy as np
import pytest

from oracle.oracle import PortOracle
from particle_simulator_b200 import FrameBuffer, default_metadata, io
from test_gpu_parity import assert_state_close

pytestmark = pytest.mark.gpu

ARGON = (3.405e-10, 1.654e-21, 12.0, 6.0)  # the metadata's second species by default is argon-like too (particle.rs:154-160)


def mixed_scene(grid, n_side=60, spacing=1.12, speeds=(20.0, 120.0), seed=61) -> FrameBuffer:
    """Two interleaved species: a lattice whose columns alternate labels, plus a block of pure species 1."""
    meta = default_metadata()
    meta["particles"][1] = ARGON
    fb = FrameBuffer(2 * n_side * n_side, meta)
    cell = 50e-9 / 64
    w, h = cell * (1 << grid[0]), cell * (1 << grid[1])
    fb.metadata["box_width"], fb.metadata["box_height"] = w, h
    io.scene_hex_square(fb, n_side, n_side, (0.27 * w, 0.5 * h), spacing, speeds[0], speeds[1], 0, seed=seed)
    p = fb.particles
    p["ty"][(np.arange(len(p)) // n_side) % 2 == 1] = 1  # every other lattice column
    io.scene_hex_square(fb, n_side, n_side, (0.73 * w, 0.5 * h), spacing, speeds[0], speeds[1], 1, seed=seed + 1)
    return fb


@pytest.mark.parametrize("grid", [(6, 6), (7, 6), (10, 10)])
def test_species_step_matches_the_oracle_extension(grid):
    from particle_simulator_b200.stepper import Stepper

    fb = mixed_scene(grid, n_side=38 if grid[0] < 10 else 120)  # on the 50 nm box the blocks reach to 1.5 nm from the walls
    port = PortOracle(grid[0], grid[1], 32)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    want = port.step_species(slots, fb.metadata, threads=4)
    want = want[want["ty"] >= 0]
    ref0 = port.step(slots, fb.metadata, threads=4)
    ref0 = ref0[ref0["ty"] >= 0]
    assert not np.array_equal(want["vx"], ref0["vx"])  # the extension does change the physics of this scene
    with Stepper(grid, fb.count, species_physics=True) as st:
        st.upload(fb)
        before = st.download().particles.copy()
        assert before.tobytes() == slots[slots["ty"] >= 0].tobytes()
        assert st.tile_stats()["float_path"] == 0
        st.step_async(1)
        st.snapshot_async()
        got = st.download().particles.copy()
    assert_state_close(got, want, before, fb.metadata, f"species step on {grid}")
    # without the switch the same scene is stepped the reference's way
    with Stepper(grid, fb.count) as st:
        st.upload(fb)
        st.step_async(1)
        st.snapshot_async()
        plain = st.download().particles.copy()
    assert_state_close(plain, ref0, before, fb.metadata, f"species-0 step on {grid}")


def test_species_switch_is_inert_when_both_species_are_the_same():
    from particle_simulator_b200.stepper import Stepper

    grid = (10, 10)
    fb = mixed_scene(grid, n_side=100)
    fb.metadata["particles"][1] = fb.metadata["particles"][0]
    outs = []
    for flag in (False, True):
        with Stepper(grid, fb.count, species_physics=flag) as st:
            st.upload(fb)
            assert st.tile_stats()["float_path"] == 1  # the fast kernel: nothing to distinguish
            st.run_frame_async()
            st.sync()
            outs.append(st.download().particles.tobytes())
    assert outs[0] == outs[1]


def test_species_frame_conserves_energy_and_keeps_labels():
    """A whole frame (101 steps, 6 re-bins) of the mixed scene: labels travel with the particles through the re-bins, the
    energy of the extension's own potential is conserved."""
    from particle_simulator_b200.stepper import Stepper

    grid = (7, 7)
    fb = mixed_scene(grid, n_side=70, spacing=1.12, speeds=(5.0, 20.0))  # 100 nm box: two 35 nm blocks
    fb.metadata["step_dt"] = 5e-15
    fb.metadata["steps_per_frame"] = 100
    with Stepper(grid, fb.count, species_physics=True) as st:
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        out = st.download()
        assert (st.steps_executed, st.rebins_executed) == (101, 6)
    assert out.count == fb.count
    assert (out.particles["ty"] == 1).sum() == (fb.particles["ty"] == 1).sum()
    ke0 = (fb.particles["vx"].astype(np.float64) ** 2 + fb.particles["vy"].astype(np.float64) ** 2).sum()
    ke1 = (out.particles["vx"].astype(np.float64) ** 2 + out.particles["vy"].astype(np.float64) ** 2).sum()
    assert 0.2 * ke0 < ke1 < 20 * ke0  # no blow-up: the unlike pairs sit near their own r0 spacing, forces stay bounded
