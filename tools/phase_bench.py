#!/usr/bin/env python
"""BASELINE.json configs[2]: the 10M-particle lattice through its heating ramp (solid -> liquid -> gas); step-kernel
time, stencil neighbours per particle and the HBM-roofline fraction in each phase.

    python tools/phase_bench.py [n_side]

Between frames the host scales the velocities (workloads.heat) and re-uploads; dt = 10 fs (the step the reference's
report calls stable for hot systems, doc/project.typ:209)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.frame import PARTICLE_MASS  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "gas":
    # the end of the ramp, built directly (the 1.2 um crystal needs nanoseconds to fill the box): the same 10M particles
    # spread over the whole box, 2.4 per cell, 350 m/s
    from particle_simulator_b200 import io
    from particle_simulator_b200.frame import FrameBuffer

    n = 10_000_000
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = fb.metadata["box_height"] = workloads.CELL_WIDTH * 2048
    fb.metadata["step_dt"] = 10e-15
    fb.metadata["steps_per_frame"] = 100
    io.scene_gas(fb, n, 2 * workloads.CELL_WIDTH, 3.4e-10, 250.0, 450.0, 0, seed=9)
    with Stepper((11, 11), n) as st:
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        st.enable_step_timing(True)
        st.run_frame_async()
        st.sync()
        ms, k = st.step_timing()
        cs = st.cell_start().astype(np.int64)
        cnt = np.diff(cs).reshape(2048, 2048)
        pad = np.pad(cnt, 1)
        stencil = sum(pad[1 + dy:2049 + dy, 1 + dx:2049 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
        nb = float((cnt * (stencil - 1)).sum() / cnt.sum())
        stats = st.tile_stats()
        kms = ms / k
        print(f"gas filling the box    neighbours/particle {nb:5.1f}  step kernel {kms:.4f} ms  "
              f"{40 * n / (kms * 1e-3) / 1e9 / 6523.7 * 100:4.1f} % of the HBM roofline  tiles {stats['tiles']} "
              f"({stats['tiles_staged']} staged), threads live/launched {stats['threads_live'] / stats['threads_launched']:.2f}")
    sys.exit(0)

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
wl = workloads.lattice(n_side, n_side + 1, (11, 11), 1.0, 1.0, 10.0, seed=3)
fb = wl.frame
fb.metadata["step_dt"] = 10e-15
fb.metadata["steps_per_frame"] = 200
n = wl.particles
peak = 6523.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def temperature(p):
    return float(PARTICLE_MASS) * float((p["vx"].astype(np.float64) ** 2 + p["vy"].astype(np.float64) ** 2).mean()) / (2 * 1.380649e-23)


def neighbours(st):
    cs = st.cell_start().astype(np.int64)
    cnt = np.diff(cs).reshape(2048, 2048)
    pad = np.pad(cnt, 1)
    stencil = sum(pad[1 + dy:2049 + dy, 1 + dx:2049 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    return float((cnt * (stencil - 1)).sum() / cnt.sum())


targets = [(0.0, "solid (as built)"), (60.0, "hot solid / melting"), (150.0, "liquid"), (600.0, "gas, expanding"), (600.0, "gas, later")]
print(f"{n} particles, 2048 x 2048 cells")
with Stepper(wl.grid_log2, n) as st:
    for t_target, label in targets:
        for _ in range(40):
            t = temperature(fb.particles)
            if t >= t_target:
                break
            workloads.heat(fb, min(1.5, max(1.05, (max(t_target, 1e-3) / max(t, 1e-6)) ** 0.5)))
            st.upload(fb)
            st.run_frame_async()
            st.sync()
            fb = st.download(fb)
        extra = 6 if "later" in label or "expanding" in label else 1
        st.upload(fb)
        for _ in range(extra):
            st.run_frame_async()
        st.sync()
        st.enable_step_timing(True)
        st.run_frame_async()
        st.sync()
        ms, k = st.step_timing()
        st.enable_step_timing(False)
        nb = neighbours(st)
        stats = st.tile_stats()
        fb = st.download(fb)
        kms = ms / k
        print(f"{label:22s} T = {temperature(fb.particles):7.1f} K  neighbours/particle {nb:5.1f}  step kernel {kms:.4f} ms  "
              f"{40 * n / (kms * 1e-3) / 1e9 / peak * 100:4.1f} % of the HBM roofline  tiles {stats['tiles']} "
              f"({stats['tiles_staged']} staged), threads live/launched {stats['threads_live'] / stats['threads_launched']:.2f}")
