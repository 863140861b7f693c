"""Generates tests/golden/*.npz from the REFERENCE's own step code (oracle/_ref/libref_6_6.so,
compiled from /root/reference by oracle/Makefile) running on Device::CpuMainThread.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The fixtures are committed; nothing at test time needs /root/reference.

Every fixture holds, for one seeded scene on the reference's fixed 64x64x16 grid:
  input          compact input frame records (what the editor would send)
  meta           the 80-byte FrameMetadata
  binned         compacted slot array right after kernel_prepare_frame (kernel.cuh:210-239)
  binned_counts  particles per cell after it
  step1          compacted slot array after ONE bucket_step from the fresh binning
  moved          compacted slot array after one bucket_move applied to step1's result
  moved_counts   particles per cell after that move
  frame_<S>      compacted result of Kernel::run_async with steps_per_frame = S from the fresh
                 binning (kernel_bucket.cuh:181-206), for each S in `frames`
  frame_steps    steps actually executed for each S (S=100 runs 101)
  diag_*         double-precision energy / momentum diagnostics of those states (oracle_diagnostics
                 over the stencil pairs of a fresh binning of the state)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.oracle import PortOracle, RefOracle  # noqa: E402
from particle_simulator_b200 import FrameBuffer, io  # noqa: E402
from particle_simulator_b200.frame import DEVICE_CPU_MAIN_THREAD, METADATA_DTYPE  # noqa: E402

SIGMA = 3.609e-10


def scene_hex2500() -> FrameBuffer:
    fb = FrameBuffer(2500)
    io.scene_hex_square(fb, 50, 50, (25e-9, 25e-9), 1.0, 5.0, 5.0, 0, seed=42)
    return fb


def scene_gas10k() -> FrameBuffer:
    """BASELINE.json configs[0]: 10k-particle gas-phase box."""
    fb = FrameBuffer(10000)
    # dt = 10 fs: with the default 50 fs a 300 K gas crosses more than half a cell between two
    # re-bins (17 steps), so two particles can approach from non-adjacent cells without ever
    # interacting and the reference itself blows up and loses particles (measured: 2259 of 10000
    # gone after 101 steps). 10 fs is the step the reference's report calls stable for leapfrog
    # (doc/project.typ:209).
    fb.metadata["step_dt"] = 10e-15
    io.scene_gas(fb, 10000, margin=2 * SIGMA, min_dist=1.1 * SIGMA, v_min=200.0, v_max=500.0, ty=0, seed=12345)
    return fb


def scene_liquid4k() -> FrameBuffer:
    fb = FrameBuffer(4096)
    io.scene_hex_square(fb, 64, 64, (25e-9, 25e-9), 1.08, 150.0, 250.0, 0, seed=1)
    return fb


def scene_wall_cursor() -> FrameBuffer:
    """Two species labels, particles pressed against the walls, an active cursor."""
    fb = FrameBuffer(3000)
    fb.metadata["cursor_pos"] = (0.3, 0.6)
    fb.metadata["cursor_size"] = 0.2
    io.scene_square(fb, 30, 30, (7e-9, 7e-9), 1.02, 50.0, 100.0, 0, seed=7)
    io.scene_square(fb, 30, 30, (43e-9, 42.5e-9), 1.0, 10.0, 150.0, 1, seed=8)
    io.scene_gas(fb, 1200, margin=1.2 * SIGMA, min_dist=1.3 * SIGMA, v_min=50.0, v_max=150.0, ty=1, seed=9)
    return fb


SCENES = {
    "hex2500": (scene_hex2500, (1, 18, 100)),
    "gas10k": (scene_gas10k, (17, 100, 1000)),
    "liquid4k": (scene_liquid4k, (100,)),
    "wall_cursor": (scene_wall_cursor, (35,)),
}


def frame_of(particles: np.ndarray, meta: np.ndarray) -> FrameBuffer:
    fb = FrameBuffer(max(len(particles), 1), np.asarray(meta).reshape(()))
    fb.set_particles(particles)
    return fb


def counts_of(slots: np.ndarray, capacity: int) -> np.ndarray:
    return (slots["ty"].reshape(-1, capacity) >= 0).sum(axis=1).astype(np.uint32)


def live(slots: np.ndarray) -> np.ndarray:
    return slots[slots["ty"] >= 0]


def main() -> None:
    ref = RefOracle(6, 6)
    port = PortOracle(6, 6, ref.capacity)
    for name, (make, frames) in SCENES.items():
        fb = make()
        fb.metadata["device"] = DEVICE_CPU_MAIN_THREAD
        n = fb.count
        out: dict[str, np.ndarray] = {
            "input": fb.particles.copy(),
            "meta": np.array(fb.metadata, dtype=METADATA_DTYPE).reshape(1),
        }

        def diag(slots: np.ndarray, meta: np.ndarray) -> np.ndarray:
            # energies are taken over the stencil pairs of a FRESH binning of the state (64 slots per
            # cell so nothing can be dropped): the pair set must not depend on how stale the
            # membership of the state happens to be, or two equal states would show different PE
            fresh = PortOracle(6, 6, 64)
            fslots, dropped = fresh.prepare(frame_of(live(slots), meta))
            assert dropped == 0
            d = fresh.diagnostics(fslots, meta)
            return np.array([d["ke"], d["pe_pair"], d["pe_wall"], d["px"], d["py"], d["live"]])

        ref.prepare(fb)
        s0 = ref.slots()
        assert counts_of(s0, ref.capacity).max() <= ref.capacity and len(live(s0)) == n
        out["binned"] = live(s0)
        out["binned_counts"] = counts_of(s0, ref.capacity)
        out["diag_binned"] = diag(s0, fb.metadata)

        ref.step()
        s1 = ref.slots()
        out["step1"] = live(s1)
        ref.move()
        s2 = ref.slots()
        assert len(live(s2)) == n, "reference lost particles in bucket_move"
        out["moved"] = live(s2)
        out["moved_counts"] = counts_of(s2, ref.capacity)

        executed = []
        for S in frames:
            fb.metadata["steps_per_frame"] = S
            ref.prepare(fb)
            ref.run_frame()
            sf = ref.slots()
            assert len(live(sf)) == n, f"reference lost particles during a {S}-step frame of {name}"
            # steps executed: replay the schedule
            steps, cd = 1, 0
            while steps < S:
                if cd <= 0:
                    cd = 15
                    steps += 1
                else:
                    cd -= 2
                    steps += 2
            executed.append(steps)
            out[f"frame_{S}"] = live(sf)
            out[f"diag_frame_{S}"] = diag(sf, fb.metadata)
        out["frames"] = np.array(frames, dtype=np.uint32)
        out["frame_steps"] = np.array(executed, dtype=np.uint32)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: n={n} max/cell={out['binned_counts'].max()} frames={frames} executed={executed} "
              f"-> {os.path.getsize(path) / 1e3:.0f} kB")
        d0, d1 = out["diag_binned"], out[f"diag_frame_{frames[-1]}"]
        print(f"   E0={d0[0] + d0[1] + d0[2]:.6e} (KE {d0[0]:.3e} PE {d0[1]:.3e} wall {d0[2]:.3e})"
              f"  E1={d1[0] + d1[1] + d1[2]:.6e} (KE {d1[0]:.3e} PE {d1[1]:.3e} wall {d1[2]:.3e})")


if __name__ == "__main__":
    main()
