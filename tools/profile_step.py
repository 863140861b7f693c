#!/usr/bin/env python
"""The smallest program that launches the step kernel on the bench workload (for ncu):
    python tools/profile_step.py [steps] [workload: solid|liquid]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
kind = sys.argv[2] if len(sys.argv) > 2 else "solid"
wl = workloads.config_10m_solid() if kind == "solid" else workloads.config_1m_liquid()
with Stepper(wl.grid_log2, wl.particles, device=0) as st:
    st.upload(wl.frame)
    print(st.tile_stats())
    st.enable_step_timing(True)
    st.step_async(steps)
    st.sync()
    ms, k = st.step_timing()
    print(f"{wl.name}: {ms / k:.4f} ms per step kernel over {k} launches")
