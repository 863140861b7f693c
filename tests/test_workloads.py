"""The synthetic workloads of BASELINE.json as frames (host side only)."""
import numpy as np

from particle_simulator_b200 import workloads
from particle_simulator_b200.slabs import slab_of


def test_clustered_mixed_scene_is_lopsided_and_reproducible():
    w = workloads.clustered_mixed((10, 10), clusters=4, side=60, gas=2000, seed=5)
    p = w.frame.particles
    assert w.particles == 4 * 60 * 60 + 2000 and len(p) == w.particles
    assert set(np.unique(p["ty"])) == {0, 1}
    assert float(w.frame.metadata["step_dt"]) == np.float32(10e-15)
    owner = slab_of(p["y"], 8, 10)
    held = np.bincount(owner, minlength=8)
    assert held.max() > 2 * held.mean()  # a row decomposition of it is badly balanced
    again = workloads.clustered_mixed((10, 10), clusters=4, side=60, gas=2000, seed=5)
    assert again.frame.tobytes() == w.frame.tobytes()
    other = workloads.clustered_mixed((10, 10), clusters=4, side=60, gas=2000, seed=6)
    assert other.frame.tobytes() != w.frame.tobytes()


def test_heat_scales_velocities_only():
    w = workloads.lattice(20, 20, (6, 6), 1.0, 1.0, 10.0, seed=2)
    before = w.frame.particles.copy()
    workloads.heat(w.frame, 1.5)
    after = w.frame.particles
    assert np.array_equal(after["x"], before["x"]) and np.array_equal(after["ty"], before["ty"])
    assert np.allclose(after["vx"], 1.5 * before["vx"]) and np.allclose(after["vy"], 1.5 * before["vy"])


def test_slab_crystal_rows_partition_the_crystal():
    geo = workloads.slab_crystal_geometry(4, per_slab=40000, rows_per_slab_log2=6, grid_x_log2=8)
    parts = [workloads.slab_crystal(r, 4, per_slab=40000, rows_per_slab_log2=6, grid_x_log2=8) for r in range(4)]
    ly = geo["grid_log2"][1]
    owned = 0
    for r, w in enumerate(parts):
        p = w.frame.particles
        owned += int((slab_of(p["y"], 4, ly) == r).sum())  # each rank is handed a little more than its own rows
    assert owned == geo["nx"] * geo["ny"]


def test_presets_round_trip_a_scene():
    """Preset::{from_frame, to_frame} and the Presets list (particle_io/src/presets.rs:84-154) through the C ABI of
    include/psim_scene.h (psim_presets_*, psim_preset_*)."""
    import pytest

    from particle_simulator_b200 import io
    from particle_simulator_b200.frame import FrameBuffer, default_metadata
    from particle_simulator_b200.presets import Presets

    fb = FrameBuffer(40)
    fb.metadata["box_width"], fb.metadata["box_height"] = 30e-9, 20e-9
    fb.metadata["particles"][1] = (3.2e-10, 0.9e-21, 11.0, 5.0)
    fb.metadata["steps_per_frame"] = 7  # not part of a preset
    io.scene_hex_square(fb, 5, 4, (15e-9, 10e-9), 1.0, 5.0, 5.0, 1, seed=1)
    lib = Presets()
    assert lib.get_presets_len() == 0
    assert lib.add_preset("droplet", fb) == 0
    p = lib.get_preset(0)
    assert p.name == "droplet" and p.particle_count == 20
    back = p.to_frame()
    assert back.count == 20 and back.particles.tobytes() == fb.particles.tobytes()
    assert float(back.metadata["box_width"]) == np.float32(30e-9) and float(back.metadata["box_height"]) == np.float32(20e-9)
    assert back.metadata["particles"].tobytes() == fb.metadata["particles"].tobytes()
    assert int(back.metadata["steps_per_frame"]) == int(default_metadata()["steps_per_frame"])  # a new frame's default
    assert lib.add_preset("again", back) == 1
    assert lib.get_presets_len() == 2 and lib.get_preset(1).name == "again"
    lib.change_preset("ignored", fb, 5)  # past the end: ignored (presets.rs:147-152)
    assert lib.get_presets_len() == 2
    empty = FrameBuffer(1)
    lib.change_preset("renamed", empty, 0)
    assert lib.get_preset(0).name == "renamed" and lib.get_preset(0).particle_count == 0
    assert lib.get_preset(0).to_frame().count == 0
    assert lib.duplicate_preset(1, "copy") == 2 and lib.get_preset(2).to_frame().particles.tobytes() == fb.particles.tobytes()
    lib.delete_preset(0)
    assert lib.get_presets_len() == 2 and lib.get_preset(0).name == "again" and lib.get_preset(1).name == "copy"
    with pytest.raises(IndexError):
        lib.delete_preset(7)
    with pytest.raises(IndexError):
        lib.get_preset(2)
    lib.close()
