"""L3 parity of the kernel the benchmark runs (step_kernel_c, particle_simulator_b200/csrc/step_float.cuh) against the
COMPILED REFERENCE on a 1024 x 1024 grid (oracle/_ref/libref_10_10.so = /root/reference's kernel.cuh with only the two
BUCKETS_*_LOG2 #defines rewritten, Device::CpuThreadPool).

One whole frame on the reference schedule (bucket_kernel_run_async, kernel_bucket.cuh:181-206: steps_per_frame = 100
executes 101 steps and 6 re-bins) -- trajectories diverge chaotically over 101 steps, so the comparison is on what a
correct integration conserves and on the thermodynamic state, with the bounds of test_gpu_parity's L3 tests on the
64 x 64 grid: particle count (exact), total energy (2e-4 of |KE| + |PE|), kinetic energy (2 %), total momentum
(2 % of m v_rms sqrt(N)).
"""
import numpy as np
import pytest

from oracle.oracle import PortOracle, RefOracle, ref_available
from particle_simulator_b200 import FrameBuffer, io
from particle_simulator_b200.frame import PARTICLE_DTYPE, PARTICLE_MASS

pytestmark = pytest.mark.gpu

CELL = 50e-9 / 64  # the reference's cell width (kernel.cuh:15-18, particle.rs:141-142)
GRID = (10, 10)


def boxed(n: int) -> FrameBuffer:
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = CELL * (1 << GRID[0])
    fb.metadata["box_height"] = CELL * (1 << GRID[1])
    return fb


def liquid_patch() -> FrameBuffer:
    """160k particles at liquid density, 100-200 m/s: the scene of test_gpu_float_path's one-step check."""
    fb = boxed(400 * 400)
    w = float(fb.metadata["box_width"])
    io.scene_hex_square(fb, 400, 400, (0.37 * w, 0.61 * w), 1.06, 100.0, 200.0, 0, seed=21)
    return fb


def corner_gas() -> FrameBuffer:
    """A 2.25-per-cell gas in one corner (staged tiles) and a thin gas elsewhere (tiles that run from global memory),
    200 m/s thermal speeds, dt = 10 fs (the step the reference is stable at for a gas)."""
    rng = np.random.default_rng(31)
    w = CELL * 1024
    gx, gy = np.meshgrid(np.arange(240), np.arange(240))
    dense = np.stack([gx.ravel(), gy.ravel()], axis=1) * 5.2e-10 + 2 * CELL + rng.uniform(-6e-11, 6e-11, (240 * 240, 2))
    tx, ty = np.meshgrid(np.arange(130), np.arange(130))
    thin = np.stack([tx.ravel(), ty.ravel()], axis=1) * 6e-9 + 4e-9 + rng.uniform(-2e-9, 2e-9, (130 * 130, 2))
    thin = thin[(thin[:, 0] > 130e-9) | (thin[:, 1] > 130e-9)]
    xy = np.concatenate([dense, thin])
    p = np.zeros(len(xy), dtype=PARTICLE_DTYPE)
    p["x"] = np.round(xy[:, 0] / w * 2**32).astype(np.uint64).astype(np.uint32)
    p["y"] = np.round(xy[:, 1] / w * 2**32).astype(np.uint64).astype(np.uint32)
    v = rng.normal(0, 200.0, (len(xy), 2))
    p["vx"], p["vy"] = v[:, 0], v[:, 1]
    fb = boxed(len(p))
    fb.set_particles(p)
    fb.metadata["step_dt"] = 10e-15
    return fb


def diagnostics(frame: FrameBuffer, meta) -> dict:
    port = PortOracle(GRID[0], GRID[1], 64)
    slots, dropped = port.prepare(frame)
    assert dropped == 0
    return port.diagnostics(slots, meta)


@pytest.mark.skipif(not ref_available(10, 10), reason="oracle/_ref/libref_10_10.so not built (make -C oracle ref)")
@pytest.mark.parametrize("scene", ["liquid_patch", "corner_gas"])
def test_whole_frame_of_step_kernel_c_vs_compiled_reference(scene):
    from particle_simulator_b200.stepper import Stepper

    fb = {"liquid_patch": liquid_patch, "corner_gas": corner_gas}[scene]()
    fb.metadata["steps_per_frame"] = 100
    n = fb.count

    # the reference: kernel_prepare_frame + Kernel::run_async + sync on its CPU thread pool
    fb.metadata["device"] = 1  # Device::CpuThreadPool
    ref = RefOracle(*GRID)
    assert ref.prepare(fb) == 1
    ref.run_frame()
    ref_out = ref.compact()
    assert ref_out.count == n  # the reference itself lost nothing (16 slots per cell, one cell per move)
    d_ref = diagnostics(ref_out, fb.metadata)

    with Stepper(GRID, n) as st:
        st.upload(fb)
        stats = st.tile_stats()
        assert stats["float_path"] == 1, stats  # the kernel under test is the one the benchmark runs
        if scene == "liquid_patch":
            assert stats["tiles_staged"] >= 0.99 * stats["tiles"]
        else:
            assert 50 < stats["tiles_staged"] < stats["tiles"]  # both the staged and the global-memory tile path
        s0, r0 = st.steps_executed, st.rebins_executed
        st.run_frame_async()
        st.sync()
        assert (st.steps_executed - s0, st.rebins_executed - r0) == (101, 6)
        assert st.tile_stats()["float_path"] == 1
        out = st.download()
    assert out.count == n
    assert np.array_equal(np.sort(out.particles["ty"]), np.sort(ref_out.particles["ty"]))
    d_gpu = diagnostics(out, fb.metadata)

    e_ref = d_ref["ke"] + d_ref["pe_pair"] + d_ref["pe_wall"]
    e_gpu = d_gpu["ke"] + d_gpu["pe_pair"] + d_gpu["pe_wall"]
    scale = abs(d_ref["ke"]) + abs(d_ref["pe_pair"]) + abs(d_ref["pe_wall"])
    print(f"{scene}: E_ref={e_ref:.6e} E_gpu={e_gpu:.6e} rel={abs(e_gpu - e_ref) / scale:.2e} "
          f"KE_ref={d_ref['ke']:.4e} KE_gpu={d_gpu['ke']:.4e} px_ref={d_ref['px']:.3e} px_gpu={d_gpu['px']:.3e}")
    assert abs(e_gpu - e_ref) <= 2e-4 * scale
    assert abs(d_gpu["ke"] - d_ref["ke"]) <= 2e-2 * abs(d_ref["ke"]) + 1e-4 * scale
    p_scale = float(PARTICLE_MASS) * np.sqrt(2 * d_ref["ke"] / float(PARTICLE_MASS) * n)  # ~ m * v_rms * sqrt(N)
    assert abs(d_gpu["px"] - d_ref["px"]) <= 2e-2 * p_scale and abs(d_gpu["py"] - d_ref["py"]) <= 2e-2 * p_scale
