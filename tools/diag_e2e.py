#!/usr/bin/env python
"""Where a step of bench.py's pipelined end-to-end loop goes: host time per call, and on the device the ingest, the frame
and the copies (CUDA events on the stepper's stream).

    python tools/diag_e2e.py [graph|plain]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from particle_simulator_b200 import workloads  # noqa: E402
from particle_simulator_b200.frame import FrameBuffer, packet_size  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

use_graph = (sys.argv[1] if len(sys.argv) > 1 else "graph") == "graph"


def pinned(n):
    t = torch.empty(packet_size(n), dtype=torch.uint8, pin_memory=True)
    return t, t.numpy()


keep = []
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:  # one slab per rank (torchrun), the weak-scaling workload of bench.py
    import torch.distributed as dist

    from particle_simulator_b200 import slabs

    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def storage(count):
        t, a = pinned(count)
        keep.append(t)
        return a

    wl = workloads.slab_crystal(rank, world, storage_factory=storage)
    cap = int(1.05 * 3162 * 3163)
    uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
    st = Stepper(wl.grid_log2, cap, device=local, slab_rank=rank, slab_count=world, ingest_capacity=wl.frame.count,
                 snapshot_buffers=2)
    st.comm_init(uid)
else:
    t, a = pinned(3162 * 3163)
    keep.append(t)
    wl = workloads.config_10m_solid(storage=a)
    cap = wl.particles
    st = Stepper(wl.grid_log2, wl.particles, device=local, snapshot_buffers=2, use_graph=use_graph)
wl.frame.metadata["steps_per_frame"] = 100
t, a = pinned(cap)
keep.append(t)
out = FrameBuffer(cap, storage=a)
stream = torch.cuda.Stream()
st.set_stream(stream.cuda_stream)
names = ["ev0", "upload_staged", "ev1", "stage_async", "run_frame_async", "ev2", "download_end", "download_begin"]
steps = 6
for rep in range(3):
    acc = dict.fromkeys(names, 0.0)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    T0 = time.perf_counter()
    st.stage_async(wl.frame)
    pending = False
    for k in range(steps):
        def tm(name, f, *args):
            t0 = time.perf_counter()
            f(*args)
            acc[name] += time.perf_counter() - t0
        tm("ev0", ev[k][0].record, stream)          # behind frame k-1 and its snapshot
        tm("upload_staged", st.upload_staged)
        tm("ev1", ev[k][1].record, stream)          # ingest k done (the call waited for it)
        if k + 1 < steps:
            tm("stage_async", st.stage_async, wl.frame)
        tm("run_frame_async", st.run_frame_async)
        tm("ev2", ev[k][2].record, stream)          # frame k + snapshot done
        if pending:
            tm("download_end", st.download_end)
        tm("download_begin", st.download_begin, out)
        pending = True
    st.download_end()
    torch.cuda.synchronize()
    tot = time.perf_counter() - T0
    ingest = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(1, steps)) / (steps - 1)
    frame = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(1, steps)) / (steps - 1)
    period = ev[1][0].elapsed_time(ev[steps - 1][0]) / (steps - 2)
    if world > 1:
        dist.barrier()
    if rank == 0:
      print(f"rep {rep} ({'graph' if use_graph else 'launch by launch'}): {1e3 * tot / steps:.2f} ms/step; host: "
          + ", ".join(f"{n} {1e3 * v / steps:.2f}" for n, v in acc.items())
          + f"; device: ingest {ingest:.2f} ms (event to event, incl. any wait for the staged copy), frame + snapshot {frame:.2f} ms, "
            f"period {period:.2f} ms", flush=True)
st.close()
if world > 1:
    dist.destroy_process_group()
