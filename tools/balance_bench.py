#!/usr/bin/env python
"""BASELINE.json configs[4] across GPUs, one process per slab: frames per second with slabs of equal numbers of cell rows
against slabs cut by psim_balance_rows. Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/balance_bench.py

Every rank builds the same seeded scene, hands the whole of it to its stepper (which keeps its own rows), runs a few
frames and times them with CUDA events on the stepper's stream; the frame time is the maximum over the ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from particle_simulator_b200 import slabs, workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper, balance_rows  # noqa: E402


def main() -> int:
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", local) if world > 1 else None)
    dev = torch.device("cuda", local)
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 1200   # droplets of side x side particles (6 of them)
    lg = int(sys.argv[3]) if len(sys.argv) > 3 else 12       # 2^lg x 2^lg cells of the reference's width
    wl = workloads.clustered_mixed((lg, lg), clusters=6, side=side, gas=side * side * 6 // 9, seed=5)
    wl.frame.metadata["steps_per_frame"] = 100
    n = wl.particles
    results = {}
    for how in ("equal rows", "balanced"):
        bounds = balance_rows(wl.frame, wl.grid_log2[1], world) if how == "balanced" else None
        uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
        st = Stepper(wl.grid_log2, n, device=local, slab_rank=rank, slab_count=world, ingest_capacity=n, bounds=bounds,
                     ghost_capacity=1 << 17, migrant_capacity=1 << 17)
        st.comm_init(uid)
        stream = torch.cuda.Stream(device=dev)
        st.set_stream(stream.cuda_stream)
        st.upload(wl.frame)
        held = st.particle_count
        st.run_frame_async()  # warm-up
        st.sync()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            t0.record()
            for _ in range(frames):
                st.run_frame_async()
            t1.record()
        st.sync()
        torch.cuda.synchronize()
        ms = slabs.reduce_scalar(dist, t0.elapsed_time(t1) / frames, "max", device=dev)
        counts = [None] * world
        dist.all_gather_object(counts, (held, st.particle_count, st.slab_info()["rows"]))
        results[how] = ms
        if rank == 0:
            start = np.array([c[0] for c in counts])
            print(f"{how:>10}: {ms:8.3f} ms per frame of 101 steps + 6 re-bins; halo mode {st.halo_mode}; rows per slab "
                  f"{[c[2] for c in counts]}; particles per slab at upload {start.tolist()} (max / mean "
                  f"{start.max() / start.mean():.2f}), after {frames + 1} frames {[c[1] for c in counts]}", flush=True)
        st.close()
    if rank == 0:
        print(f"{n} particles, {world} slabs: balanced boundaries are {results['equal rows'] / results['balanced']:.2f}x "
              f"faster than equal rows", flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        code = main()
    except BaseException:
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    sys.stdout.flush()
    os._exit(code)
