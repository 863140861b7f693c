"""One rank of the NCCL slab test (launched by torchrun from tests/test_gpu_nccl.py, one process per GPU).

Every rank owns one slab of the 1M-particle melting liquid (BASELINE.json configs[1]); halo exchange and
migration go through NCCL send/recv (psim_comm_init). After every frame the slabs' snapshots, concatenated
in rank order, must be byte-identical to the single-slab run of the same scene (rank 0 runs it too)."""
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
    dist.init_process_group("gloo")
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    wl = workloads.config_1m_liquid()
    wl.frame.metadata["steps_per_frame"] = 52  # 52 steps, 3 re-bins
    n = wl.particles
    uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128)
    # the lattice splits evenly at first and melts across the boundaries: leave room for the imbalance
    bounds = [int(b) for b in os.environ["PSIM_TEST_BOUNDS"].split(",")] if os.environ.get("PSIM_TEST_BOUNDS") else None
    st = Stepper(wl.grid_log2, int(0.75 * n) if world > 1 else n, device=local, slab_rank=rank, slab_count=world,
                 ingest_capacity=n, bounds=bounds)
    st.comm_init(uid)
    want_mode = {"push": 2, "nccl": 1}.get(os.environ.get("PSIM_EXPECT_HALO", ""), 0)
    if rank == 0:
        print(f"halo mode {st.halo_mode} (1: send/recv after every step, 2: pushed by the step kernel)", flush=True)
    if want_mode and world > 1 and st.halo_mode != want_mode:
        raise RuntimeError(f"rank {rank}: halo mode {st.halo_mode}, expected {want_mode}")
    single = Stepper(wl.grid_log2, n, device=local) if rank == 0 else None
    st.upload(wl.frame)  # the whole scene: the slab keeps its own rows
    if single:
        single.upload(wl.frame)
    ok = True
    counts_seen = set()
    meta_change = os.environ.get("PSIM_TEST_META_CHANGE") == "1"
    for frame in range(frames + 1):
        if frame >= 2 and meta_change:
            # a header-only metadata update between frames: another sigma (frame 2), then another box (frame 3) --
            # the neighbour records change scale while the ghost rows are being pushed (team_refresh_stale_records)
            m = wl.frame.metadata.copy()
            m["particles"][0]["sigma"] = np.float32(3.609e-10 * (1.0 + 0.01 * frame))
            if frame >= 3:
                m["box_width"] = np.float32(float(m["box_width"]) * 1.005)
                m["box_height"] = np.float32(float(m["box_height"]) * 1.005)
            st.set_metadata(m)
            if single:
                single.set_metadata(m)
        if frame:
            st.run_frame_async()
            st.sync()
            if single:
                single.run_frame_async()
                single.sync()
        mine = st.download().particles.tobytes()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, 0)
        if rank == 0:
            want = single.download().particles.tobytes()
            got = b"".join(gathered)
            same = got == want
            counts = tuple(len(g) // 20 for g in gathered)
            counts_seen.add(counts)
            print(f"frame {frame}: slabs hold {counts}, identical to the single-slab run: {same}", flush=True)
            ok = ok and same and sum(counts) == n
    # the pipelined calls on slabs (copies go out in pieces behind the re-bins): scene in, one frame, snapshot out, twice
    from particle_simulator_b200.frame import FrameBuffer

    if not meta_change:
        piped = Stepper(wl.grid_log2, int(0.75 * n) if world > 1 else n, device=local, slab_rank=rank, slab_count=world,
                        ingest_capacity=n, bounds=bounds, snapshot_buffers=2)
        piped.comm_init(slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128))
        outs = [FrameBuffer(int(0.75 * n) if world > 1 else n) for _ in range(2)]
        piped.stage_async(wl.frame)
        for k in range(2):
            piped.upload_staged()
            piped.stage_async(wl.frame)
            piped.run_frame_async()
            if k:
                piped.download_end()
            piped.download_begin(outs[k])
        piped.download_end()
        if single:
            single.upload(wl.frame)
            single.run_frame_async()
            single.sync()
        for k in range(2):
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(outs[k].particles.tobytes(), gathered, 0)
            if rank == 0:
                same = b"".join(gathered) == single.download().particles.tobytes()
                print(f"pipelined frame {k}: identical to the single-slab run: {same}", flush=True)
                ok = ok and same
        piped.close()
    if rank == 0 and world > 1 and len(counts_seen) < 2:
        print("no particle ever changed slab: the migration path was not exercised", flush=True)
        ok = False
    flag = torch.tensor([1 if ok else 0])
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
