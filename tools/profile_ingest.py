#!/usr/bin/env python
"""The smallest program that ingests the bench scene twice and re-bins it once (for an ncu launch list):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_ingest.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

wl = workloads.config_10m_solid()
with Stepper(wl.grid_log2, wl.particles, device=0) as st:
    st.upload(wl.frame)
    st.upload(wl.frame)
    st.step_async(2)
    st.rebin_async()
    st.sync()
