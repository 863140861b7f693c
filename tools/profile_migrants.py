#!/usr/bin/env python
"""Re-bins of a hot gas cut into slabs on ONE GPU (in-process slab group): the scene with the most migrants per re-bin.
Under `ncu --metrics gpu__time_duration.sum -k regex:migrant` it lists the time of the migrant kernels.

    python tools/profile_migrants.py [slabs] [particles]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import io, workloads  # noqa: E402
from particle_simulator_b200.frame import FrameBuffer  # noqa: E402
from particle_simulator_b200.stepper import SlabGroup  # noqa: E402

slabs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
fb = FrameBuffer(n)
fb.metadata["box_width"] = fb.metadata["box_height"] = workloads.CELL_WIDTH * 2048
fb.metadata["step_dt"] = 10e-15
fb.metadata["steps_per_frame"] = 34
io.scene_gas(fb, n, 2 * workloads.CELL_WIDTH, 3.4e-10, 250.0, 450.0, 0, seed=9)
with SlabGroup((11, 11), slabs, int(1.1 * n / slabs), ingest_capacity=n, migrant_capacity=131072, device=0) as gr:
    gr.upload(fb)
    gr.run_frame_async()
    gr.sync()
    sent = sum(s.migrants_sent for s in gr.slabs)
    rebins = gr.slabs[0].rebins_executed
    print(f"{slabs} slabs, {n} particles: {sent} migrants over {rebins} re-bins = {sent / max(rebins, 1) / slabs:.0f} per slab and re-bin")
