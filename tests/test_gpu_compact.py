"""DataStructure::CompactArray (the reference's all-pairs mode, cuda_simulator/src/kernel_compact.cuh:4-92): every
particle interacts with every other one, particles keep their input order, a frame is exactly steps_per_frame steps.
Checked against the reference's own compact_step_kernel (compiled into oracle/_ref)."""
import numpy as np
import pytest

from conftest import frame_from
from oracle.oracle import RefOracle, ref_available
from particle_simulator_b200 import FrameBuffer, default_metadata, io
from particle_simulator_b200.frame import COMPACT_ARRAY, PARTICLE_DTYPE
from test_gpu_parity import assert_state_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_available(6, 6), reason="oracle/_ref/libref_6_6.so has not been built")]


def compact_scene(n_lattice=20, gas=400, seed=3) -> FrameBuffer:
    meta = default_metadata()
    meta["data_structure"] = COMPACT_ARRAY
    fb = FrameBuffer(n_lattice * n_lattice + gas, meta)
    io.scene_hex_square(fb, n_lattice, n_lattice, (12e-9, 30e-9), 1.05, 50.0, 150.0, 0, seed=seed)
    io.scene_gas(fb, gas, 1.5e-9, 1.5e-9, 100.0, 300.0, 1, seed=seed + 1)  # all over the box: separations > box / 2
    return fb


@pytest.mark.parametrize("mie", [None, (3.404e-10, 117.84 * 1.380649e-23, 12.085, 6.0), (3.3e-10, 1.1e-21, 11.3, 6.5)])
def test_one_all_pairs_step_vs_the_reference(mie):
    from particle_simulator_b200.stepper import Stepper

    fb = compact_scene()
    if mie:
        fb.metadata["particles"][0] = mie
    ref = RefOracle(6, 6)
    want = ref.compact_step(fb.particles, fb.metadata)
    with Stepper((6, 6), 4096) as st:
        st.upload(fb)
        before = st.download().particles.copy()
        assert before.tobytes() == fb.particles.tobytes()  # input order, no binning
        st.step_async(1)
        st.snapshot_async()
        got = st.download().particles.copy()
    assert_state_close(got, want, before, fb.metadata, f"all pairs {mie}")


def test_null_records_are_dropped_in_place_and_frames_run_exactly_s_steps():
    from particle_simulator_b200.stepper import Stepper

    fb = compact_scene(12, 100)
    p = fb.particles.copy()
    p["ty"][::7] = -1
    fb.set_particles(p)
    live = p[p["ty"] >= 0]
    ref = RefOracle(6, 6)
    with Stepper((6, 6), 4096) as st:
        for s_per_frame, steps in ((5, 5), (4, 4), (1, 1), (0, 2)):  # kernel_compact.cuh:78-92
            fb.metadata["steps_per_frame"] = s_per_frame
            st.upload(fb)
            assert st.particle_count == len(live)
            assert st.download().particles.tobytes() == live.tobytes()
            s0 = st.steps_executed
            st.run_frame_async()
            st.sync()
            assert st.steps_executed - s0 == steps and st.rebins_executed == 0
        # the 2-step frame, step by step through the reference
        want = ref.compact_step(ref.compact_step(live, fb.metadata), fb.metadata)
        got = st.download().particles
        assert np.array_equal(got["ty"], want["ty"])
        dx = np.abs((got["x"].astype(np.int64) - want["x"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        assert dx.max() <= 8
        # a grid scene afterwards switches back
        fb.metadata["data_structure"] = 1
        fb.metadata["steps_per_frame"] = 18
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        assert st.rebins_executed == 1
        # a header-only update cannot switch the layout of a scene that is already there (include/psim_b200.h)
        from particle_simulator_b200.stepper import PsimError

        flipped = fb.metadata.copy()
        flipped["data_structure"] = COMPACT_ARRAY
        with pytest.raises(PsimError, match="data_structure"):
            st.set_metadata(flipped)
        assert int(st.get_metadata()["data_structure"]) == 1
        st.set_metadata(fb.metadata)  # the same layout: accepted


def test_all_pairs_cannot_be_decomposed_into_slabs():
    from particle_simulator_b200.stepper import PsimError, SlabGroup

    fb = compact_scene(8, 10)
    with SlabGroup((6, 6), 2, 4096, ingest_capacity=4096) as gr:
        with pytest.raises(PsimError, match="cannot be decomposed"):
            gr.upload(fb)
