"""BASELINE.json's other configurations as parity / property cases (the bench line is configs[2], 10 M particles):
configs[4] mixed-species clustered scene (heterogeneous density, load imbalance, migration) and the heating ramp of
configs[2] (solid -> liquid -> gas: the neighbour density the kernel sees changes by a factor of two)."""
import numpy as np
import pytest

from conftest import frame_from
from oracle.oracle import PortOracle
from particle_simulator_b200 import FrameBuffer, io, workloads
from particle_simulator_b200.frame import PARTICLE_MASS
from test_gpu_parity import assert_state_close

pytestmark = pytest.mark.gpu


def test_species_label_is_carried_and_species0_parameters_are_used():
    """Two species with different Mie parameters in the metadata: the reference steps both with particles[0]
    (kernel_bucket.cuh:52); the label only travels. Checked against the oracle on the reference's own grid."""
    fb = FrameBuffer(2 * 40 * 40 + 300)
    fb.metadata["particles"][1] = (3.2e-10, 0.9e-21, 11.0, 5.0)  # must not matter
    io.scene_hex_square(fb, 40, 40, (14e-9, 20e-9), 1.05, 100.0, 200.0, 0, seed=1)
    io.scene_hex_square(fb, 40, 40, (34e-9, 30e-9), 1.05, 100.0, 200.0, 1, seed=2)
    io.scene_gas(fb, 300, 1.5e-9, 1.2e-9, 200.0, 500.0, 1, seed=3)
    from particle_simulator_b200.stepper import Stepper

    port = PortOracle(6, 6, 16)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    before = slots[slots["ty"] >= 0]
    want = port.step(slots, fb.metadata)
    want = want[want["ty"] >= 0]
    with Stepper((6, 6), 8192) as st:
        st.upload(fb)
        st.step_async(1)
        st.snapshot_async()
        got = st.download().particles.copy()
    assert_state_close(got, want, before, fb.metadata, "mixed species")
    assert (got["ty"] == 1).sum() == 40 * 40 + 300


def test_mixed_species_clusters_in_8_slabs_imbalance_and_migration():
    """configs[4] at test size on a 1024 x 1024 grid: 8 slabs hold very different numbers of particles, the fast gas
    keeps crossing slab boundaries, and the result is bit-identical to the single slab, labels included."""
    from particle_simulator_b200.stepper import SlabGroup, Stepper

    w = workloads.clustered_mixed((10, 10), clusters=6, side=150, gas=20000, seed=5)
    w.frame.metadata["steps_per_frame"] = 150  # 150 steps, 9 re-bins
    n = w.particles
    species1 = int((w.frame.particles["ty"] == 1).sum())
    with Stepper(w.grid_log2, n) as st, SlabGroup(w.grid_log2, 8, n, ingest_capacity=n) as gr:
        st.upload(w.frame)
        gr.upload(w.frame)
        held = []
        for _ in range(3):
            st.run_frame_async()
            gr.run_frame_async()
            st.sync()
            gr.sync()
            a, b = st.download().particles, gr.download().particles
            assert a.tobytes() == b.tobytes()
            assert len(a) == n and int((a["ty"] == 1).sum()) == species1
            held.append([s.particle_count for s in gr.slabs])
        assert st.tile_stats()["float_path"] == 1
    held = np.array(held)
    assert held.max() / held.mean() > 2.0, held          # a row decomposition of this scene is badly balanced
    assert (np.abs(np.diff(held, axis=0)).sum(axis=1) >= 10).all(), held  # particles changed slab in every frame


def test_heating_ramp_melts_and_evaporates_the_crystal():
    """configs[2]'s ramp at test size: every frame the host scales the velocities and re-uploads. The crystal melts and
    expands; nothing is lost, the kernel keeps staging its tiles while the density drops."""
    from particle_simulator_b200.stepper import Stepper

    w = workloads.lattice(220, 220, (10, 10), 1.0, 1.0, 10.0, seed=2)
    w.frame.metadata["steps_per_frame"] = 200
    w.frame.metadata["step_dt"] = 10e-15  # the step the reference's report calls stable (doc/project.typ:209)
    n = w.particles
    fb = w.frame

    def temperature(p):  # 2-D: m <v^2> / (2 k_B)
        return float(PARTICLE_MASS) * float((p["vx"].astype(np.float64) ** 2 + p["vy"].astype(np.float64) ** 2).mean()) / (2 * 1.380649e-23)

    def extent(p):
        return float(p["x"].max() - p["x"].min()) / 2**32

    t0, e0 = temperature(fb.particles), extent(fb.particles)
    stats = []
    with Stepper(w.grid_log2, n) as st:
        for frame in range(40):
            st.upload(fb)
            stats.append(st.tile_stats())
            st.run_frame_async()
            st.sync()
            fb = st.download()
            assert fb.count == n
            assert np.isfinite(fb.particles["vx"]).all() and np.isfinite(fb.particles["vy"]).all()
            if temperature(fb.particles) < 600.0:
                workloads.heat(fb, 1.6)
    t1, e1 = temperature(fb.particles), extent(fb.particles)
    print(f"T {t0:.2f} K -> {t1:.1f} K, extent {e0:.3f} -> {e1:.3f} of the box, "
          f"columns per tile {stats[0]['max_columns']} -> {stats[-1]['max_columns']}")
    assert t1 > 100 * t0 and t1 > 100.0
    assert e1 > 1.3 * e0                                        # it evaporated
    assert stats[-1]["max_columns"] > stats[0]["max_columns"]  # fewer particles per cell: tiles span more columns
    # the solid and the liquid are staged in shared memory; the thin vapour around them runs from global memory
    assert all(s["float_path"] == 1 for s in stats) and stats[0]["tiles_staged"] >= 0.95 * stats[0]["tiles"]
    assert stats[-1]["tiles_staged"] >= 0.5 * stats[-1]["tiles"]
    print(f"staged tiles {stats[0]['tiles_staged']}/{stats[0]['tiles']} -> {stats[-1]['tiles_staged']}/{stats[-1]['tiles']}")
