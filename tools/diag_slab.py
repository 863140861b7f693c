#!/usr/bin/env python
"""Step-kernel time of the weak-scaling crystal: one slab vs the in-process slab group on ONE GPU.

    python tools/diag_slab.py [world] [scale]

Prints the average step-kernel time per slab (CUDA events around every launch) so that a slow slab
path can be told apart from a slow transport.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.stepper import SlabGroup, Stepper  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    geo = workloads.slab_crystal_geometry(world)
    wl = workloads.lattice(geo["nx"], geo["ny"], geo["grid_log2"], name="crystal")
    wl.frame.metadata["steps_per_frame"] = 35
    n = wl.particles
    print(f"crystal {geo['nx']}x{geo['ny']} = {n} particles, grid 2^{geo['grid_log2']}", flush=True)
    with Stepper(wl.grid_log2, n, device=0) as st:
        st.upload(wl.frame)
        st.run_frame_async()
        st.sync()
        st.enable_step_timing(True)
        st.run_frame_async()
        st.sync()
        ms, k = st.step_timing()
        print(f"single slab: {ms / k:.4f} ms per step kernel over {k} launches ({n} particles)", flush=True)
    with SlabGroup(wl.grid_log2, world, int(1.05 * n / world), ingest_capacity=n, device=0) as gr:
        gr.upload(wl.frame)
        gr.run_frame_async()
        gr.sync()
        for s in gr.slabs:
            s.enable_step_timing(True)
        gr.run_frame_async()
        gr.sync()
        for r, s in enumerate(gr.slabs):
            ms, k = s.step_timing()
            print(f"slab {r}: {ms / k:.4f} ms per step kernel over {k} launches ({s.particle_count} particles, "
                  f"{s.slab_info()})", flush=True)


if __name__ == "__main__":
    main()
