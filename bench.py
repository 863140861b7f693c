#!/usr/bin/env python
"""bench.py -- particle-updates/s of the per-timestep update at 10M particles (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A bench "step" is one FRAME of the hot path: metadata.steps_per_frame = 100 leapfrog steps on the
reference's step / re-bin schedule (101 steps and 6 re-bins are executed, kernel_bucket.cuh:181-206).
  value        particle-updates/s with the state resident in HBM (frames chained on the device)
  e2e          the same metric through the reference-facing C ABI with HOST frames: every step
               uploads the scene (psim_upload_frame, H2D from pinned memory + binning), runs a frame
               and downloads the compacted result (psim_download_frame, D2H)
  roofline     step kernel only: 40 B per particle-update (20 B record in + 20 B out)
               over the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline the reference's own CPU implementation (oracle/_ref, compiled from /root/reference)
               on this box's host cores, on a bounded sample of the same workload
`--impl reference` times that CPU implementation alone and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ALGO_BYTES_PER_UPDATE = 40
STEPS_PER_FRAME = 100


def measured_hbm_peak() -> tuple[float, str]:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic() -> float | None:
    """dram bytes per launch of the step kernel from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(REPO, "profiles", "step_kernel_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, name in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index: int) -> str:
    """Run this rank on the CPU cores next to its GPU, so that the page-locked frames it allocates (first touch) and
    the copies to and from them stay on the GPU's NUMA node. With several ranks per host the end-to-end leg is
    bound by host memory traffic; ranks left floating over the sockets halve it. Best effort: returns what it did."""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        domain, rest = bus.split(":", 1)
        dev = f"{domain[-4:]}:{rest}"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read())
        if node < 0:
            return "single NUMA node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: none of its cores is available"
        os.sched_setaffinity(0, cpus)
        return f"node {node}, {len(cpus)} cores"
    except Exception as exc:  # no sysfs entry, no nvidia-smi, a container without the topology ...
        return f"not bound ({type(exc).__name__})"


def pinned_storage(nbytes: int):
    import torch

    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    return t, t.numpy()


def run_reference(args, rank: int, world: int) -> None:
    """The reference's CPU implementation of the path (Device::CpuThreadPool, all host threads) from
    oracle/_ref on the same 10M-particle workload; a step = one frame of `ref_steps` leapfrog steps."""
    if rank != 0:
        return
    from oracle.oracle import RefOracle, ref_available
    from particle_simulator_b200 import workloads
    from particle_simulator_b200.frame import DEVICE_CPU_THREAD_POOL

    if not ref_available(11, 11):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_11_11.so has not been built"}))
        return
    wl = workloads.config_10m_solid()
    ref = RefOracle(11, 11)
    wl.frame.metadata["device"] = DEVICE_CPU_THREAD_POOL
    wl.frame.metadata["steps_per_frame"] = args.ref_steps
    ref.prepare(wl.frame)
    executed = schedule_steps(args.ref_steps)
    for _ in range(args.warmup):
        ref.run_frame()
    t = 0.0
    for _ in range(args.steps):
        t += ref.run_frame()
    n = wl.particles
    value = n * executed * args.steps / t
    sample = (f"{args.steps} frames of {executed} leapfrog steps (steps_per_frame={args.ref_steps}) on the full "
              f"{n}-particle scene, Device::CpuThreadPool")
    line = {
        "impl": "reference", "metric": "particle-updates/sec at 10M particles", "value": value,
        "unit": "particle-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "particles": n,
                   "steps_per_frame": args.ref_steps, "leapfrog_steps_per_bench_step": executed},
        "cpu_baseline": {"value": value, "unit": "particle-updates/s", "cores": ref.hardware_threads,
                         "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def schedule_steps(S: int) -> int:
    """Steps the reference schedule executes for steps_per_frame = S (kernel_bucket.cuh:181-206)."""
    steps, cd = 1, 0
    while steps < S:
        if cd <= 0:
            cd, steps = 15, steps + 1
        else:
            cd, steps = cd - 2, steps + 2
    return steps


def cpu_baseline(wl, budget_steps: int) -> dict:
    """Reference CPU implementation on a bounded sample, for the `cpu_baseline` object of our own line."""
    from oracle.oracle import RefOracle, ref_available
    from particle_simulator_b200.frame import DEVICE_CPU_THREAD_POOL

    lx, ly = wl.grid_log2
    if not ref_available(lx, ly):
        return {"value": None, "unit": "particle-updates/s", "cores": 0, "kind": "reference",
                "sample": f"oracle/_ref/libref_{lx}_{ly}.so not built"}
    ref = RefOracle(lx, ly)
    fb = wl.frame.copy()
    fb.metadata["device"] = DEVICE_CPU_THREAD_POOL
    fb.metadata["steps_per_frame"] = budget_steps
    ref.prepare(fb)
    executed = schedule_steps(budget_steps)
    t = ref.run_frame()
    return {"value": wl.particles * executed / t, "unit": "particle-updates/s", "cores": ref.hardware_threads,
            "kind": "reference",
            "sample": f"1 frame of {executed} leapfrog steps on the full {wl.particles}-particle scene, "
                      f"Device::CpuThreadPool ({t:.1f} s)"}


def reference_cuda_baseline(wl, budget_steps: int) -> dict:
    """The reference's EXISTING CUDA backend (bucket_step_gpu / bucket_move_gpu, kernel_bucket.cuh:96-110, compiled
    for sm_100a into oracle/_ref -- its own build ships sm_86 / sm_89 SASS only) on the same B200 and the same scene:
    Device::Gpu, gpu_threads_per_block_log2 = 7, Kernel::run_async + sync timed (no copies). A reported baseline."""
    from oracle.oracle import RefOracle, ref_available
    from particle_simulator_b200.frame import DEVICE_GPU

    lx, ly = wl.grid_log2
    if not ref_available(lx, ly):
        return {"value": None, "unit": "particle-updates/s", "kind": "reference-cuda",
                "sample": f"oracle/_ref/libref_{lx}_{ly}.so not built"}
    ref = RefOracle(lx, ly)
    if ref.gpu_count == 0:
        return {"value": None, "unit": "particle-updates/s", "kind": "reference-cuda", "sample": "the reference found no GPU"}
    fb = wl.frame.copy()
    fb.metadata["device"] = DEVICE_GPU
    fb.metadata["steps_per_frame"] = budget_steps
    used = ref.prepare(fb)
    executed = schedule_steps(budget_steps)
    ref.run_frame()  # warm-up
    frames = 2
    t = sum(ref.run_frame() for _ in range(frames))
    return {"value": wl.particles * executed * frames / t, "unit": "particle-updates/s", "kind": "reference-cuda",
            "ms_per_leapfrog_step": 1e3 * t / (frames * executed),
            "sample": f"{frames} frames of {executed} leapfrog steps on the full {wl.particles}-particle scene, the "
                      f"reference's bucket_step_gpu / bucket_move_gpu rebuilt for sm_100a, device field {used}, "
                      f"{ref.slot_count} slots of {ref.capacity} per cell ({t:.2f} s)"}


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist

    from particle_simulator_b200 import slabs, workloads
    from particle_simulator_b200.frame import FrameBuffer, packet_size
    from particle_simulator_b200.stepper import Stepper

    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(x: float, op: str) -> float:
        return slabs.reduce_scalar(dist, x, op, device=dev) if world > 1 else x

    # the scene lives in page-locked host memory: it is what the e2e leg uploads every step
    keep = []

    def pinned_frame_storage(count: int) -> np.ndarray:
        t, a = pinned_storage(packet_size(count))
        keep.append(t)
        return a

    stream = torch.cuda.Stream()
    strong = args.scaling == "strong"
    if strong:
        # strong scaling (BASELINE.json configs[3]): ONE crystal of --total-particles on a fixed 8192 x 4096 grid
        # (box 6.4 x 3.2 um), cut into `world` slabs of 4096 / world cell rows
        rows_log2 = 12 - (world.bit_length() - 1)
        per_slab = args.total_particles // world
        wl = workloads.slab_crystal(rank, world, storage_factory=pinned_frame_storage, per_slab=per_slab,
                                    rows_per_slab_log2=rows_log2, grid_x_log2=13)
        st = Stepper(wl.grid_log2, int(1.05 * per_slab) + 65536, device=local_rank, slab_rank=rank, slab_count=world,
                     ingest_capacity=wl.frame.count, snapshot_buffers=2)
        if world > 1:
            uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
            st.comm_init(uid)
    elif world == 1:
        wl = workloads.config_10m_solid(storage=pinned_frame_storage(3162 * 3163))
        st = Stepper(wl.grid_log2, wl.particles, device=local_rank, snapshot_buffers=2)
    else:
        # weak scaling: one crystal across `world` slabs of 2048 cell rows, ~10M particles per slab; rank r
        # steps slab r, halo rows and migrants travel over NCCL send/recv (NVLink)
        wl = workloads.slab_crystal(rank, world, storage_factory=pinned_frame_storage)
        uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
        st = Stepper(wl.grid_log2, int(1.05 * 3162 * 3163), device=local_rank, slab_rank=rank, slab_count=world,
                     ingest_capacity=wl.frame.count, snapshot_buffers=2)
        st.comm_init(uid)
    halo = {0: "none (single slab)", 1: "ncclSend/ncclRecv after every step",
            2: "pushed by the step kernel over NVLink peer memory (CUDA IPC), epoch flags"}[st.halo_mode]
    wl.frame.metadata["steps_per_frame"] = STEPS_PER_FRAME
    executed = schedule_steps(STEPS_PER_FRAME)
    st.set_stream(stream.cuda_stream)

    # ---- device-resident leg -------------------------------------------------------------------
    st.upload(wl.frame)
    n_local = st.particle_count
    n = int(over_ranks(float(n_local), "sum"))  # particles of the whole job
    cap_local = st.max_particles  # particles migrate between slabs: a slab's count is not constant
    out = FrameBuffer(cap_local, storage=pinned_frame_storage(cap_local))
    for _ in range(args.warmup):
        st.run_frame_async()
    st.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    st.enable_step_timing(not args.no_step_timing)
    launches0 = st.kernel_launches
    steps0 = st.steps_executed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t_host = time.perf_counter()
    for _ in range(args.steps):
        st.run_frame_async()
    t_host = time.perf_counter() - t_host  # host time to enqueue the timed frames (launch-bound if ~ms)
    e1.record(stream)
    st.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    step_ms, step_launches = st.step_timing()
    st.enable_step_timing(False)
    launches = st.kernel_launches - launches0
    steps_done = st.steps_executed - steps0
    assert steps_done == executed * args.steps
    ms = over_ranks(ms, "max")
    value = n * steps_done / (ms * 1e-3)

    # ---- end-to-end leg: host frame in, host frame out, every step -------------------------------
    # Pipelined through the public API: every step uploads its scene from page-locked host memory
    # (psim_stage_frame_async + psim_upload_staged) and downloads its result (psim_download_frame_begin / _end); the
    # copies of steps k+1 and k-1 run on their own streams while step k's frame is computed.
    def e2e_pipelined(steps: int) -> float:
        barrier()
        t0 = time.perf_counter()
        st.stage_async(wl.frame)
        pending = False
        for k in range(steps):
            st.upload_staged()
            if k + 1 < steps:
                st.stage_async(wl.frame)
            st.run_frame_async()
            if pending:
                st.download_end()
            st.download_begin(out)
            pending = True
        st.download_end()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    # ... and one call after the other (upload, frame, download), nothing overlapped, for comparison
    def e2e_synchronous(steps: int) -> float:
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            st.upload(wl.frame)
            st.run_frame_async()
            st.download(out)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_pipelined(min(args.warmup, 2))
    t_e2e = over_ranks(e2e_pipelined(args.steps), "max")
    t_e2e_sync = over_ranks(e2e_synchronous(args.steps), "max")
    assert int(over_ranks(float(out.count), "sum")) == n  # nothing lost, whatever slab holds it now
    e2e_value = n * executed * args.steps / t_e2e
    h2d = int(over_ranks(float(packet_size(wl.frame.count)), "sum"))
    d2h = int(over_ranks(float(packet_size(out.count)), "sum"))
    kernel_ms_max = over_ranks(step_ms / max(step_launches, 1), "max")
    launches = int(over_ranks(float(launches), "sum"))

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        kernel_ms = step_ms / max(step_launches, 1)
        achieved = ALGO_BYTES_PER_UPDATE * n_local / (kernel_ms * 1e-3) / 1e9  # per GPU (rank 0's slab)
        line = {
            "metric": "particle-updates/sec at 10M particles", "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "description": wl.description, "particles": n,
                       "particles_rank0": n_local,
                       "steps_per_frame": STEPS_PER_FRAME, "leapfrog_steps_per_bench_step": executed,
                       "rebins_per_bench_step": 6, "schedule": "reference (kernel_bucket.cuh:181-206)",
                       "l2": "state (10M x 20 B x 2 buffers = 400 MB per GPU) is larger than the 126 MB L2; "
                             "no flush",
                       "decomposition": "single slab" if world == 1 else
                       f"{world} slabs of {st.slab_info()['rows']} cell rows, one per GPU; halo (boundary rows' positions) every step: "
                       f"{halo}; per re-bin: migrants + boundary-row cell counts + fresh ghost rows by ncclSend/ncclRecv",
                       "step_kernel": st.tile_stats()},
            "roofline": {"bound": "hbm", "kernel": "step_kernel (fused 3x3-cell force + kick + drift)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": recorded_traffic(),
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * n_local,
                         "kernel_ms": kernel_ms, "kernel_ms_max_over_ranks": kernel_ms_max, "kernel_launches_timed": step_launches,
                         "kernel_share_of_step": step_ms / ms},
            "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "how": "every step: scene uploaded from page-locked host memory, binned, one frame, snapshot "
                           "downloaded; the copies of steps k+1 / k-1 overlap step k's frame "
                           "(psim_stage_frame_async, psim_upload_staged, psim_download_frame_begin/_end)",
                    "synchronous_ms_per_step": 1e3 * t_e2e_sync / args.steps},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": 1e3 * t_host / args.steps,
            "host_numa_binding_rank0": numa,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline and not strong:
            st.close()  # the reference allocates its own slot arrays on this GPU
            line["cpu_baseline"] = cpu_baseline(wl, args.cpu_steps)
            line["reference_cuda_baseline"] = reference_cuda_baseline(wl, args.cpu_steps)
        print(json.dumps(line))
    st.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--ref-steps", type=int, default=2, help="--impl reference: steps_per_frame of one bench step")
    ap.add_argument("--cpu-steps", type=int, default=10, help="cpu_baseline: steps_per_frame of the sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-step-timing", action="store_true", help="no CUDA events around the step-kernel launches")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak (default, the metric's configuration): 10M particles per GPU; strong: --total-particles "
                         "in one fixed 8192 x 4096-cell box cut into --gpus slabs (BASELINE.json configs[3])")
    ap.add_argument("--total-particles", type=int, default=100_000_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
