#!/usr/bin/env python
"""The smallest program that runs re-bins on the bench workload (for ncu): upload, a few steps, three re-bins."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

wl = workloads.config_10m_solid()
with Stepper(wl.grid_log2, wl.particles, device=0) as st:
    st.upload(wl.frame)
    for _ in range(3):
        st.step_async(4)
        st.rebin_async()
    st.sync()
    print("re-bins:", st.rebins_executed)
