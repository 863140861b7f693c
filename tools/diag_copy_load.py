#!/usr/bin/env python
"""Does host<->device copy traffic slow the frames down? (the 8-GPU end-to-end question of VERDICT r01 item 5)

    python -m torch.distributed.run --nproc-per-node N ... tools/diag_copy_load.py [slab|replica]

Every rank runs device-resident frames of the bench workload, first alone, then while two side streams copy 200 MB
buffers to and from page-locked host memory back to back (the traffic of the end-to-end loop, with no dependency on
the frames at all).
  slab:    the ranks are the slabs of one crystal (7 host round trips per frame: 6 re-bin commits + none at ingest)
  replica: every rank steps its own single slab, frames replayed as CUDA graphs (no host round trip inside a frame)
"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from particle_simulator_b200 import slabs, workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "slab"
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if mode == "slab" and world > 1:
    wl = workloads.slab_crystal(rank, world)
    uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
    st = Stepper(wl.grid_log2, int(1.05 * 3162 * 3163), device=local, slab_rank=rank, slab_count=world,
                 ingest_capacity=wl.frame.count)
    st.comm_init(uid)
else:
    wl = workloads.config_10m_solid()
    st = Stepper(wl.grid_log2, wl.particles, device=local, use_graph=True)
wl.frame.metadata["steps_per_frame"] = 100
stream = torch.cuda.Stream()
st.set_stream(stream.cuda_stream)
st.upload(wl.frame)
st.run_frame_async()
st.sync()

nbytes = 200 * 1000 * 1000
h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
stop = threading.Event()
copied = [0]


def copy_load():
    torch.cuda.set_device(local)
    while not stop.is_set():
        with torch.cuda.stream(s_in):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_b, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        copied[0] += 1


def frames(k: int) -> float:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(k):
        st.run_frame_async()
    e1.record(stream)
    st.sync()
    t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


alone = frames(5)
th = threading.Thread(target=copy_load, daemon=True)
th.start()
time.sleep(0.3)
c0, t0 = copied[0], time.perf_counter()
loaded = frames(5)
rate = (copied[0] - c0) * nbytes / (time.perf_counter() - t0) / 1e9
stop.set()
th.join()
r = torch.tensor([rate], dtype=torch.float64, device=dev)
dist.all_reduce(r, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{mode}, {world} ranks: frame {alone:.2f} ms alone, {loaded:.2f} ms while every rank copies 200 MB each way back to back "
          f"(slowest rank: {float(r.item()):.1f} GB/s each way)", flush=True)
st.close()
dist.destroy_process_group()
