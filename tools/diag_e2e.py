#!/usr/bin/env python
"""Host-side timeline of bench.py's pipelined e2e loop (where does a step's wall time go?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from particle_simulator_b200 import workloads
from particle_simulator_b200.frame import FrameBuffer, packet_size
from particle_simulator_b200.stepper import Stepper

def pinned(n):
    t = torch.empty(packet_size(n), dtype=torch.uint8, pin_memory=True); return t, t.numpy()
keep = []
t, a = pinned(3162 * 3163); keep.append(t)
wl = workloads.config_10m_solid(storage=a)
wl.frame.metadata["steps_per_frame"] = 100
t, a = pinned(wl.particles); keep.append(t)
out = FrameBuffer(wl.particles, storage=a)
st = Stepper(wl.grid_log2, wl.particles, device=0, snapshot_buffers=2)
names = ["upload_staged", "stage_async", "run_frame_async", "download_end", "download_begin"]
for rep in range(2):
    acc = dict.fromkeys(names, 0.0)
    torch.cuda.synchronize(); T0 = time.perf_counter()
    st.stage_async(wl.frame); pending = False
    for k in range(4):
        def tm(name, f, *args):
            t0 = time.perf_counter(); f(*args); acc[name] += time.perf_counter() - t0
        tm("upload_staged", st.upload_staged)
        if k + 1 < 4: tm("stage_async", st.stage_async, wl.frame)
        tm("run_frame_async", st.run_frame_async)
        if pending: tm("download_end", st.download_end)
        tm("download_begin", st.download_begin, out); pending = True
    st.download_end(); torch.cuda.synchronize()
    tot = time.perf_counter() - T0
    print(f"rep {rep}: {1e3*tot/4:.2f} ms/step; " + ", ".join(f"{n} {1e3*v/4:.2f}" for n, v in acc.items()))
