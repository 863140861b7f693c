"""Parity of the CUDA stepper (through the C ABI of include/psim_b200.h) with the reference.

Three levels (BASELINE.json north_star):
  L1 bit-exact  : cell of every particle, per-cell counts, particle order after ingest and re-bin
  L2 tolerance  : state after exactly one step from a fresh binning -- velocities within
                  1e-5 * max(|v|, a_max*dt) (a_max = largest Mie attraction / mass), positions within
                  2 fixed-point LSB + 1e-5 of the step's displacement
  L3 statistical: energy and momentum after whole frames (trajectories diverge chaotically)
Checked against the golden vectors generated from the reference's own compiled code and against the
oracle restatement (itself pinned to those vectors by tests/test_oracle.py).
"""
import numpy as np
import pytest

from conftest import frame_from
from oracle.oracle import PortOracle
from particle_simulator_b200 import FrameBuffer, default_metadata, io
from particle_simulator_b200.frame import PARTICLE_DTYPE, PARTICLE_MASS

pytestmark = pytest.mark.gpu

SCENES = ["hex2500", "gas10k", "liquid4k", "wall_cursor"]
V_RTOL = 1e-5   # north_star: "within a stated fp32 relative tolerance (e.g. 1e-5)"
X_LSB = 2       # fixed-point LSBs of slack on positions (the final roundf)


@pytest.fixture(scope="module")
def Stepper():
    from particle_simulator_b200.stepper import Stepper as S

    return S


def a_max_dt(meta) -> float:
    """Velocity change per step under the largest attractive Mie force."""
    p = meta["particles"][0]
    n, m, sigma, eps = (float(p[k]) for k in ("n", "m", "sigma", "epsilon"))
    C = n / (n - m) * (n / m) ** (m / (n - m))
    r = sigma * (((n + 1) * n) / ((m + 1) * m)) ** (1 / (n - m))
    f = abs(C * eps * (m * (sigma / r) ** m - n * (sigma / r) ** n) / r)
    return f / float(PARTICLE_MASS) * float(meta["step_dt"])


def assert_state_close(got: np.ndarray, want: np.ndarray, before: np.ndarray, meta, what: str,
                       max_pair_force: np.ndarray | None = None):
    """L2 comparison of two particle arrays in the same order. `max_pair_force` (newtons, per particle):
    the largest single pair force acting on the particle; where it exceeds the Mie attraction scale the
    velocity tolerance is relative to it (a net force is a near-cancelling sum of such terms)."""
    assert len(got) == len(want), what
    assert np.array_equal(got["ty"], want["ty"]), what + ": species labels / order differ"
    scale = a_max_dt(meta)
    if max_pair_force is not None:
        scale = np.maximum(scale, max_pair_force.astype(np.float64) / float(PARTICLE_MASS) * float(meta["step_dt"]))
    for c in ("vx", "vy"):
        tol = V_RTOL * np.maximum(np.abs(want[c].astype(np.float64)), scale)
        err = np.abs(got[c].astype(np.float64) - want[c].astype(np.float64))
        worst = int(np.argmax(err / tol))
        assert (err <= tol).all(), f"{what}: {c}[{worst}] got {got[c][worst]!r} want {want[c][worst]!r} " \
                                   f"err {err[worst]:.3e} tol {tol[worst]:.3e}"
    for c in ("x", "y"):
        disp = np.abs((want[c].astype(np.int64) - before[c].astype(np.int64) + 2**31) % 2**32 - 2**31)
        box = float(meta["box_width" if c == "x" else "box_height"])
        tol = X_LSB + V_RTOL * np.maximum(disp, scale * float(meta["step_dt"]) / box * 2**32)
        err = np.abs((got[c].astype(np.int64) - want[c].astype(np.int64) + 2**31) % 2**32 - 2**31)
        worst = int(np.argmax(err - tol))
        assert (err <= tol).all(), f"{what}: {c}[{worst}] err {err[worst]} LSB, tol {tol[worst]:.1f}"


def cells_of(p: np.ndarray, lx: int, ly: int) -> np.ndarray:
    cx = (p["x"] >> np.uint32(32 - lx)).astype(np.int64)
    cy = (p["y"] >> np.uint32(32 - ly)).astype(np.int64)
    return cx + (cy << lx)


def stable_sort_by_cell(p: np.ndarray, lx: int, ly: int) -> np.ndarray:
    live = p[p["ty"] >= 0]
    return live[np.argsort(cells_of(live, lx, ly), kind="stable")]


# ------------------------------------------------------------------------------------------------
# L1: binning
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", SCENES)
def test_ingest_binning_bitexact_vs_reference(name, golden, Stepper):
    g = golden(name)
    fb = frame_from(g["input"], g["meta"][0])
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        assert st.particle_count == len(g["input"])
        cs = st.cell_start()
        assert cs[0] == 0 and cs[-1] == len(g["input"])
        assert np.array_equal(np.diff(cs), g["binned_counts"])
        out = st.download()
        assert out.is_valid() and out.metadata.tobytes() == fb.metadata.tobytes()
        assert out.particles.tobytes() == g["binned"].tobytes()


@pytest.mark.parametrize("name", SCENES)
def test_rebin_of_reference_state_bitexact(name, golden, Stepper):
    """bucket_move == stable re-sort by the new cell: feeding the reference's post-step state through
    the GPU sort must give the reference's post-move state, counts and order."""
    g = golden(name)
    fb = frame_from(g["step1"], g["meta"][0])
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        assert np.array_equal(np.diff(st.cell_start()), g["moved_counts"])
        assert st.download().particles.tobytes() == g["moved"].tobytes()


@pytest.mark.parametrize("name", SCENES)
def test_live_rebin_is_the_stable_sort_of_the_live_state(name, golden, Stepper):
    """The re-bin kernels working on the device-resident structure of arrays (not the ingest path):
    step a few times, snapshot, re-bin, snapshot; the second must be the stable sort of the first."""
    g = golden(name)
    fb = frame_from(g["input"], g["meta"][0])
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        st.step_async(17)
        st.snapshot_async()
        before = st.download().particles.copy()
        st.rebin_async()
        st.snapshot_async()
        after = st.download().particles.copy()
        want = stable_sort_by_cell(before, 6, 6)
        assert after.tobytes() == want.tobytes()
        counts = np.bincount(cells_of(want, 6, 6), minlength=4096)
        assert np.array_equal(np.diff(st.cell_start()), counts)
        # idempotence: sorting a sorted state changes nothing
        st.rebin_async()
        st.snapshot_async()
        assert st.download().particles.tobytes() == after.tobytes()
        assert st.rebins_executed == 2 and st.steps_executed == 17


def test_ingest_skips_null_particles_and_keeps_input_order(Stepper):
    rng = np.random.default_rng(5)
    n = 5000
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    p["x"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    p["y"] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    p["vx"] = np.arange(n)  # tags the input order
    p["ty"] = rng.integers(0, 2, n)
    p["ty"][rng.random(n) < 0.25] = -1
    fb = frame_from(p, default_metadata())
    with Stepper((5, 7), 8192) as st:  # non-square grid
        st.upload(fb)
        want = stable_sort_by_cell(p, 5, 7)
        assert st.particle_count == len(want)
        assert st.download().particles.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------
# L2: one step
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", SCENES)
def test_single_step_vs_reference_golden(name, golden, Stepper):
    g = golden(name)
    fb = frame_from(g["input"], g["meta"][0])
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        st.step_async(1)
        st.snapshot_async()
        got = st.download().particles
        assert_state_close(got, g["step1"], g["binned"], fb.metadata, name)


def gpu_and_port_one_step(Stepper, fb, lx, ly, capacity=16, steps=1):
    port = PortOracle(lx, ly, capacity)
    slots, dropped = port.prepare(fb)
    assert dropped == 0
    before = slots[slots["ty"] >= 0]
    cur = slots
    for _ in range(steps):
        cur = port.step(cur, fb.metadata, threads=8)
    want = cur[cur["ty"] >= 0]
    with Stepper((lx, ly), max(fb.count, 1)) as st:
        st.upload(fb)
        assert st.download().particles.tobytes() == before.tobytes()
        st.step_async(steps)
        st.snapshot_async()
        got = st.download().particles.copy()
    return got, want, before


def test_single_step_midsize_liquid_vs_oracle(Stepper):
    # 200k particles, 256x256 cells (same cell width as the reference grid => box 200 nm)
    fb = FrameBuffer(448 * 448)
    fb.metadata["box_width"] = 200e-9
    fb.metadata["box_height"] = 200e-9
    io.scene_hex_square(fb, 448, 448, (100e-9, 100e-9), 1.07, 100.0, 200.0, 0, seed=11)
    got, want, before = gpu_and_port_one_step(Stepper, fb, 8, 8)
    assert_state_close(got, want, before, fb.metadata, "liquid200k")


@pytest.mark.parametrize("mie", [(3.404e-10, 117.84 * 1.380649e-23, 12.085, 6.0),   # argon: other n
                                 (3.609e-10, 1.46e-21, 12.0, 6.0),                  # Lennard-Jones 12-6
                                 (3.3e-10, 1.1e-21, 11.3, 6.5),                     # fractional m
                                 (3.609e-10, 1.46e-21, 9.0, 4.0),                   # odd n, small m
                                 (3.609e-10, 1.46e-21, 16.4, 6.0),                  # q^9 * 2^z through MUFU.EX2
                                 (3.609e-10, 1.46e-21, 14.3, 6.0),                  # q^8, rest too big for the cubic
                                 (3.609e-10, 1.46e-21, 10.0, 6.0),                  # q^6, no rest
                                 (3.609e-10, 1.46e-21, 18.2, 6.0),                  # q^10
                                 (3.609e-10, 1.46e-21, 8.1, 6.0),                   # q^5 and a small rest
                                 (3.609e-10, 1.46e-21, 24.0, 6.0)])                 # beyond the compile-time powers
def test_single_step_other_mie_parameters_vs_oracle(mie, Stepper):
    fb = FrameBuffer(60 * 60)
    fb.metadata["particles"][0] = mie
    io.scene_hex_square(fb, 60, 60, (25e-9, 25e-9), 1.05, 50.0, 150.0, 0, seed=12)
    got, want, before = gpu_and_port_one_step(Stepper, fb, 6, 6)
    assert_state_close(got, want, before, fb.metadata, f"mie{mie}")


def test_single_step_nonsquare_grid_and_box_vs_oracle(Stepper):
    fb = FrameBuffer(120 * 50)
    fb.metadata["box_width"] = 100e-9   # 128 x 32 cells of 7.8125e-10 m
    fb.metadata["box_height"] = 25e-9
    io.scene_hex_square(fb, 120, 50, (50e-9, 12.5e-9), 1.06, 50.0, 150.0, 0, seed=13)
    got, want, before = gpu_and_port_one_step(Stepper, fb, 7, 5)
    assert_state_close(got, want, before, fb.metadata, "nonsquare")


def test_single_step_sparse_scene_takes_the_global_memory_path(Stepper):
    # 300 particles in 4096 cells: a tile of 128 particles spans ~1700 cells, more than the staging
    # buffers hold, so the kernel reads cell_start / positions from global memory instead
    fb = FrameBuffer(300)
    io.scene_gas(fb, 300, margin=1e-9, min_dist=4e-10, v_min=100, v_max=300, seed=14)
    got, want, before = gpu_and_port_one_step(Stepper, fb, 6, 6, steps=3)
    assert_state_close(got, want, before, fb.metadata, "sparse")


def test_single_step_crowded_cells_beyond_reference_capacity(Stepper):
    # a compressed lattice puts ~25 particles in a cell; the reference (16 slots) would drop some
    # (kernel_bucket.cuh:31), the oracle restatement is run with 64 slots per cell instead
    fb = FrameBuffer(40 * 40)
    fb.metadata["box_width"] = 12.5e-9  # 16 x 16 cells
    fb.metadata["box_height"] = 12.5e-9
    fb.metadata["step_dt"] = 1e-15
    io.scene_square(fb, 40, 40, (6.25e-9, 6.25e-9), 0.4, 0.0, 10.0, 0, seed=15)
    port = PortOracle(4, 4, 64)
    slots, dropped = port.prepare(fb)
    assert dropped == 0 and (slots["ty"].reshape(-1, 64) >= 0).sum(axis=1).max() > 16
    got, want, before = gpu_and_port_one_step(Stepper, fb, 4, 4, capacity=64)
    _, _, max_pair = port.forces(slots, fb.metadata)
    assert_state_close(got, want, before, fb.metadata, "crowded", max_pair[slots["ty"] >= 0])


def test_cursor_and_wall_forces_single_particle(Stepper):
    # no pairs at all: only the cursor (kernel_bucket.cuh:54-67) and wall (particle.cuh:125-144) terms
    meta = default_metadata()
    meta["cursor_pos"] = (0.5, 0.5)
    meta["cursor_size"] = 0.4
    for (fx, fy) in [(0.45, 0.52), (0.02, 0.985), (0.99, 0.012), (0.5, 0.5), (0.3, 0.0065)]:
        p = np.zeros(1, dtype=PARTICLE_DTYPE)
        p["x"], p["y"] = int(fx * 2**32), int(fy * 2**32)
        p["vx"], p["vy"] = 3.0, -4.0
        fb = frame_from(p, meta)
        got, want, before = gpu_and_port_one_step(Stepper, fb, 6, 6)
        assert_state_close(got, want, before, meta, f"single@{fx},{fy}")


def test_empty_scene(Stepper):
    fb = FrameBuffer(1)
    with Stepper((6, 6), 16) as st:
        st.upload(fb)
        assert st.particle_count == 0
        st.run_frame_async()
        st.sync()
        out = st.download()
        assert out.count == 0 and out.is_valid()
        assert st.steps_executed == 101


def test_capacity_error_and_state_errors(Stepper):
    from particle_simulator_b200.stepper import PsimError

    fb = FrameBuffer(100)
    io.scene_square(fb, 10, 10, (25e-9, 25e-9))
    with Stepper((6, 6), 50) as st:
        with pytest.raises(PsimError, match="max_particles"):
            st.upload(fb)
        with pytest.raises(PsimError, match="no scene"):
            st.run_frame_async()
    with pytest.raises(PsimError):
        Stepper((1, 6), 16)


# ------------------------------------------------------------------------------------------------
# the frame schedule and L3
# ------------------------------------------------------------------------------------------------

def test_reference_schedule_step_and_rebin_counts(golden, Stepper):
    g = golden("hex2500")
    fb = frame_from(g["input"], g["meta"][0])
    table = {1: (1, 0), 2: (2, 1), 3: (4, 1), 16: (16, 1), 17: (18, 1), 18: (18, 1), 19: (19, 2), 100: (101, 6)}
    with Stepper((6, 6), 4096) as st:
        for S, (steps, moves) in table.items():
            fb.metadata["steps_per_frame"] = S
            st.upload(fb)
            s0, r0 = st.steps_executed, st.rebins_executed
            st.run_frame_async()
            st.sync()
            assert (st.steps_executed - s0, st.rebins_executed - r0) == (steps, moves), S


def test_native_schedule_runs_exactly_n_steps(golden, Stepper):
    from particle_simulator_b200.stepper import SCHEDULE_NATIVE

    g = golden("hex2500")
    fb = frame_from(g["input"], g["meta"][0])
    fb.metadata["steps_per_frame"] = 25
    with Stepper((6, 6), 4096, schedule=SCHEDULE_NATIVE, rebin_every=10) as st:
        st.upload(fb)
        st.run_frame_async()
        st.run_frame_async()
        st.sync()
        assert st.steps_executed == 50
        assert st.rebins_executed == 4  # before steps 10, 20, 30, 40 (the ingest binned step 0)
        assert st.download().count == 2500


def diagnostics_of(particles: np.ndarray, meta, lx=6, ly=6, cap=64):
    port = PortOracle(lx, ly, cap)
    slots, dropped = port.prepare(frame_from(particles, meta))
    assert dropped == 0
    return port.diagnostics(slots, meta)


@pytest.mark.parametrize("name,S", [("hex2500", 18), ("hex2500", 100), ("liquid4k", 100), ("gas10k", 100),
                                    ("wall_cursor", 35)])
def test_frame_energy_and_momentum_vs_reference(name, S, golden, Stepper):
    g = golden(name)
    fb = frame_from(g["input"], g["meta"][0])
    fb.metadata["steps_per_frame"] = S
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        out = st.download()
        executed = int(g["frame_steps"][list(g["frames"]).index(S)])
        assert st.steps_executed == executed
    assert out.count == len(g["input"])  # nothing lost
    d = diagnostics_of(out.particles, fb.metadata)
    ke, pe, pw, px, py, _ = g[f"diag_frame_{S}"]
    e_ref, e_gpu = ke + pe + pw, d["ke"] + d["pe_pair"] + d["pe_wall"]
    scale = abs(ke) + abs(pe) + abs(pw)
    print(f"{name} S={S}: E_ref={e_ref:.6e} E_gpu={e_gpu:.6e} rel={abs(e_gpu - e_ref) / scale:.2e} "
          f"KE_ref={ke:.4e} KE_gpu={d['ke']:.4e}")
    assert abs(e_gpu - e_ref) <= 2e-4 * scale
    assert abs(d["ke"] - ke) <= 2e-2 * abs(ke) + 1e-4 * scale
    p_scale = float(PARTICLE_MASS) * np.sqrt(2 * ke / float(PARTICLE_MASS) * len(g["input"]))  # ~ m * v_rms * sqrt(N)
    assert abs(d["px"] - px) <= 2e-2 * p_scale and abs(d["py"] - py) <= 2e-2 * p_scale
    # the lattice scenes have not had time to diverge: particle by particle they still agree
    if name == "hex2500":
        want = g[f"frame_{S}"]
        assert np.array_equal(out.particles["ty"], want["ty"])
        dx = np.abs((out.particles["x"].astype(np.int64) - want["x"].astype(np.int64) + 2**31) % 2**32 - 2**31)
        assert dx.max() < 2**32 * 1e-6  # < 1e-6 of the box (0.05 pm)


def test_config1_gas_1000_steps_energy_drift(golden, Stepper):
    """BASELINE.json configs[0] on the GPU: 10k-particle gas, 1000 leapfrog steps (59 re-bins)."""
    g = golden("gas10k")
    fb = frame_from(g["input"], g["meta"][0])
    fb.metadata["steps_per_frame"] = 1000
    with Stepper((6, 6), 65536) as st:
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        assert (st.steps_executed, st.rebins_executed) == (1000, 59)
        out = st.download()
    assert out.count == 10000
    d = diagnostics_of(out.particles, fb.metadata)
    ke0, pe0, pw0 = g["diag_binned"][:3]
    ke1, pe1, pw1 = g["diag_frame_1000"][:3]
    e0, e_ref, e_gpu = ke0 + pe0 + pw0, ke1 + pe1 + pw1, d["ke"] + d["pe_pair"] + d["pe_wall"]
    scale = abs(ke0) + abs(pe0)
    drift_ref, drift_gpu = (e_ref - e0) / scale, (e_gpu - e0) / scale
    print(f"config1: drift_ref={drift_ref:.3e} drift_gpu={drift_gpu:.3e} KE_ref={ke1:.4e} KE_gpu={d['ke']:.4e}")
    # chaotic system: not the same trajectory, but the same thermodynamic state and the same drift scale
    assert abs(drift_gpu) <= 2 * abs(drift_ref) + 2e-3
    assert abs(d["ke"] - ke1) <= 0.03 * ke1
    p_scale = float(PARTICLE_MASS) * np.sqrt(2 * ke1 / float(PARTICLE_MASS) * 10000)
    assert abs(d["px"] - g["diag_frame_1000"][3]) <= 0.1 * p_scale
    assert abs(d["py"] - g["diag_frame_1000"][4]) <= 0.1 * p_scale


def test_pair_forces_conserve_momentum(Stepper):
    # far from the walls and with the cursor off, one step changes total momentum only by rounding:
    # pair forces are antisymmetric (f_dist is exactly odd, particle.cuh:41-47)
    fb = FrameBuffer(64 * 64)
    io.scene_hex_square(fb, 64, 64, (25e-9, 25e-9), 1.1, 0.0, 0.0, 0, seed=16)
    with Stepper((6, 6), 4096) as st:
        st.upload(fb)
        st.step_async(1)
        st.snapshot_async()
        out = st.download().particles
    dv = np.abs(out["vx"].astype(np.float64)).sum() + np.abs(out["vy"].astype(np.float64)).sum()
    assert dv > 1.0  # forces did act (stretched lattice)
    assert abs(out["vx"].astype(np.float64).sum()) < 1e-5 * dv
    assert abs(out["vy"].astype(np.float64).sum()) < 1e-5 * dv
