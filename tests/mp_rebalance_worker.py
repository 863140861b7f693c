"""One rank of the cross-process rebalancing test (launched by torchrun from tests/test_gpu_nccl.py, one process per GPU).

The clustered scene of BASELINE.json configs[4] at test size, cut into slabs of EQUAL rows (badly balanced), runs a
frame; then slabs.rebalance_across_ranks moves the boundaries to where the particles are (all-reduced row histogram,
all-to-all of the records, steppers re-created) and two more frames run. Rank 0 also runs the single slab, which
restarts its schedule at the same point by taking its own state back in. The slabs' snapshots, concatenated in rank
order, must stay byte-identical to the single slab's, and the imbalance must drop below 1.1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from particle_simulator_b200 import slabs, workloads  # noqa: E402
from particle_simulator_b200.stepper import Stepper  # noqa: E402


def main() -> int:
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = workloads.clustered_mixed((10, 10), clusters=6, side=150, gas=20000, seed=5)
    wl.frame.metadata["steps_per_frame"] = 35
    n = wl.particles

    def make(bounds):
        uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
        st = Stepper(wl.grid_log2, n, device=local, slab_rank=rank, slab_count=world, ingest_capacity=n, bounds=bounds,
                     ghost_capacity=1 << 15, migrant_capacity=1 << 15)
        st.comm_init(uid)
        return st

    st = make(None)
    single = Stepper(wl.grid_log2, n, device=local) if rank == 0 else None
    st.upload(wl.frame)
    if single:
        single.upload(wl.frame)
    ok = True
    imbalance = []

    def compare(label: str) -> None:
        nonlocal ok
        mine = st.download().particles
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = len(mine)
        dist.all_reduce(counts)
        counts = counts.cpu().tolist()
        pad = torch.zeros(max(counts) * 20, dtype=torch.uint8, device=dev)
        pad[:len(mine) * 20] = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        imbalance.append(max(counts) / (sum(counts) / world))
        if rank == 0:
            got = b"".join(g[:c * 20].cpu().numpy().tobytes() for g, c in zip(gathered, counts))
            same = got == single.download().particles.tobytes()
            print(f"{label}: slabs hold {counts} (max / mean {imbalance[-1]:.2f}), identical to the single-slab run: {same}",
                  flush=True)
            ok = ok and same and sum(counts) == n

    def frame() -> None:
        st.run_frame_async()
        st.sync()
        if single:
            single.run_frame_async()
            single.sync()

    frame()
    compare("equal rows, frame 1")
    st, bounds = slabs.rebalance_across_ranks(dist, st, make, device=dev)
    if single:
        single.snapshot_async()
        single.upload(single.download())
    compare(f"rebalanced to rows {bounds}")
    for k in range(2):
        frame()
        compare(f"balanced, frame {k + 2}")
    if rank == 0 and not (imbalance[0] > 1.3 and max(imbalance[1:]) < 1.1):
        print(f"imbalance before / after: {imbalance}", flush=True)
        ok = False
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    st.close()
    if single:
        single.close()
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    try:
        code = main()
    except BaseException:  # a rank that dies must not leave its peers waiting in a collective
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    sys.stdout.flush()
    os._exit(code)
