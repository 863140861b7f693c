"""Pins the oracle: the plain-C restatement (oracle/psim_oracle.c) against the golden vectors that
were generated from the REFERENCE's own compiled code (tests/golden/make_golden.py), and -- in the
build container, where oracle/_ref exists -- against that compiled reference directly."""
import numpy as np
import pytest

from conftest import frame_from
from oracle.oracle import PortOracle, RefOracle, ref_available

SCENES = ["hex2500", "gas10k", "liquid4k", "wall_cursor"]


def live(slots):
    return slots[slots["ty"] >= 0]


def counts(slots, cap=16):
    return (slots["ty"].reshape(-1, cap) >= 0).sum(axis=1).astype(np.uint32)


@pytest.mark.parametrize("name", SCENES)
def test_port_binning_step_move_bitexact_vs_golden(name, golden):
    g = golden(name)
    port = PortOracle(6, 6, 16)
    fb = frame_from(g["input"], g["meta"][0])
    s0, dropped = port.prepare(fb)
    assert dropped == 0
    assert np.array_equal(counts(s0), g["binned_counts"])
    assert live(s0).tobytes() == g["binned"].tobytes()
    s1 = port.step(s0, fb.metadata)
    assert live(s1).tobytes() == g["step1"].tobytes()
    s2, alive = port.move(s1)
    assert alive == len(g["input"])
    assert np.array_equal(counts(s2), g["moved_counts"])
    assert live(s2).tobytes() == g["moved"].tobytes()


@pytest.mark.parametrize("name", SCENES)
def test_port_frames_bitexact_vs_golden(name, golden):
    g = golden(name)
    port = PortOracle(6, 6, 16)
    fb = frame_from(g["input"], g["meta"][0])
    s0, _ = port.prepare(fb)
    for S, executed in zip(g["frames"], g["frame_steps"]):
        if S > 200:
            continue  # the long run has its own test below
        meta = fb.metadata.copy()
        meta["steps_per_frame"] = S
        out, steps, moves = port.run_frame(s0, meta, threads=4)
        assert steps == executed
        assert live(out).tobytes() == g[f"frame_{S}"].tobytes()


def test_port_1000_steps_config1_bitexact_vs_golden(golden):
    """BASELINE.json configs[0]: 10k-particle gas box, 1000 leapfrog steps on the CPU."""
    g = golden("gas10k")
    port = PortOracle(6, 6, 16)
    fb = frame_from(g["input"], g["meta"][0])
    s0, _ = port.prepare(fb)
    meta = fb.metadata.copy()
    meta["steps_per_frame"] = 1000
    out, steps, moves = port.run_frame(s0, meta, threads=8)
    assert (steps, moves) == (1000, 59)
    assert live(out).tobytes() == g["frame_1000"].tobytes()
    d = port.diagnostics(out, meta)
    want = g["diag_frame_1000"]
    assert np.allclose([d["ke"], d["pe_pair"], d["pe_wall"], d["px"], d["py"]], want[:5], rtol=1e-12)


def test_schedule_step_counts():
    # steps executed / moves for steps_per_frame S (kernel_bucket.cuh:181-206)
    port = PortOracle(2, 2, 4)
    from particle_simulator_b200 import default_metadata

    slots = np.zeros(port.slot_count, dtype=port.prepare(frame_from(np.zeros(0, dtype=[("x", "<u4"), ("y", "<u4"),
                     ("vx", "<f4"), ("vy", "<f4"), ("ty", "<i4")]), default_metadata()))[0].dtype)
    slots["ty"] = -1
    table = {1: (1, 0), 2: (2, 1), 3: (4, 1), 16: (16, 1), 17: (18, 1), 18: (18, 1), 19: (19, 2), 100: (101, 6)}
    for S, want in table.items():
        meta = default_metadata()
        meta["steps_per_frame"] = S
        _, steps, moves = port.run_frame(slots, meta)
        assert (steps, moves) == want, S


def test_scalar_physics_constants():
    from particle_simulator_b200 import default_metadata

    port = PortOracle(6, 6, 16)
    meta = default_metadata()
    assert abs(port.params_C(meta) - 3.283061) < 1e-6
    r0 = 4.01084e-10
    f_scale = abs(port.f_force(meta, 4.41047e-10))
    assert abs(port.f_force(meta, r0)) < 1e-4 * f_scale  # force vanishes at r0
    assert abs(f_scale - 1.0493e-11) < 2e-15  # max attraction


needs_ref = pytest.mark.skipif(not ref_available(6, 6), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("name", SCENES)
def test_port_matches_compiled_reference(name, golden):
    g = golden(name)
    ref = RefOracle(6, 6)
    port = PortOracle(6, 6, ref.capacity)
    fb = frame_from(g["input"], g["meta"][0])
    fb.metadata["device"] = 2  # CpuMainThread
    fb.metadata["steps_per_frame"] = 20
    ref.prepare(fb)
    s_port, _ = port.prepare(fb)
    a = ref.slots()
    assert np.array_equal(a["ty"], s_port["ty"]) and live(a).tobytes() == live(s_port).tobytes()
    ref.run_frame()
    out, steps, moves = port.run_frame(s_port, fb.metadata)
    b = ref.slots()
    assert np.array_equal(b["ty"], out["ty"]) and live(b).tobytes() == live(out).tobytes()
    assert ref.params_C(fb.metadata) == port.params_C(fb.metadata)
    for r in (3.5e-10, 4.0e-10, 4.4e-10, 7.8e-10):
        assert ref.f_force(fb.metadata, r) == port.f_force(fb.metadata, r)


@needs_ref
def test_reference_threadpool_equals_main_thread(golden):
    # CPU executors agree bit for bit (every slot is independent)
    g = golden("hex2500")
    ref = RefOracle(6, 6)
    outs = []
    for dev in (2, 1):
        fb = frame_from(g["input"], g["meta"][0])
        fb.metadata["device"] = dev
        fb.metadata["steps_per_frame"] = 18
        ref.prepare(fb)
        ref.run_frame()
        outs.append(ref.compact().particles.tobytes())
    assert len(outs[0]) == 2500 * 20 and outs[0] == outs[1]


@needs_ref
def test_all_pairs_kernel_agrees_inside_one_stencil(golden):
    """kernel_compact.cuh:4-34 is the same physics over all pairs: for a cluster that fits inside one
    3x3 stencil it must give the same forces up to summation order."""
    from particle_simulator_b200 import FrameBuffer, io

    ref = RefOracle(6, 6)
    port = PortOracle(6, 6, 16)
    fb = FrameBuffer(9)
    # a 3x3 patch (pitch r0 = 0.51 cells) centred on a cell corner: it touches cells 32..33 only, so
    # every pair is inside the stencil
    cw = 50e-9 / 64
    io.scene_square(fb, 3, 3, (33.0 * cw, 33.0 * cw), 1.0, 20.0, 40.0, 0, seed=4)
    want = ref.compact_step(fb.particles, fb.metadata)
    s0, _ = port.prepare(fb)
    got = live(port.step(s0, fb.metadata))

    # the 9 particles spread over several cells, so the two outputs are in different orders:
    # match them by position (they are ~1e6 fixed-point units apart, a step moves them far less)
    def by_position(p):
        return p[np.lexsort((p["x"] >> 20, p["y"] >> 20))]

    w, g = by_position(want), by_position(got)
    assert np.array_equal(w["ty"], g["ty"])
    assert np.abs(w["x"].astype(np.int64) - g["x"].astype(np.int64)).max() <= 2
    assert np.abs(w["y"].astype(np.int64) - g["y"].astype(np.int64)).max() <= 2
    assert np.allclose(w["vx"], g["vx"], rtol=1e-5, atol=1e-4)
    assert np.allclose(w["vy"], g["vy"], rtol=1e-5, atol=1e-4)


def test_species_extension_reduces_to_the_reference_step():
    """oracle_step_species (the spec of PsimConfig.species_physics, an extension): with identical species, or with only
    species-0 labels, it IS the pinned reference step; with distinct species it differs, and pair forces stay
    antisymmetric (total momentum changes only by rounding away from the walls)."""
    from particle_simulator_b200 import FrameBuffer, default_metadata, io

    meta = default_metadata()
    meta["particles"][1] = (3.405e-10, 1.654e-21, 12.0, 6.0)
    fb = FrameBuffer(2 * 30 * 30, meta)
    io.scene_hex_square(fb, 30, 30, (18e-9, 25e-9), 1.1, 10.0, 60.0, 0, seed=71)
    io.scene_hex_square(fb, 30, 30, (32e-9, 25e-9), 1.1, 10.0, 60.0, 1, seed=72)
    port = PortOracle(6, 6, 16)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    ref = port.step(slots, fb.metadata)
    ext = port.step_species(slots, fb.metadata)
    live = slots["ty"] >= 0
    s0 = live & (slots["ty"] == 0)
    far0 = s0 & (slots["x"] < np.uint32(0.45 * 2**32))  # species-0 particles out of reach of the species-1 block
    assert ext[far0].tobytes() == ref[far0].tobytes()
    assert not np.array_equal(ext[live & (slots["ty"] == 1)]["vx"], ref[live & (slots["ty"] == 1)]["vx"])
    same = fb.metadata.copy()
    same["particles"][1] = same["particles"][0]
    assert port.step_species(slots, same).tobytes() == port.step(slots, same).tobytes()
    dv = ext["vx"][live].astype(np.float64) - slots["vx"][live].astype(np.float64)
    assert abs(dv.sum()) < 1e-4 * np.abs(dv).sum()
