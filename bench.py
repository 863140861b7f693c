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


def recorded_traffic() -> tuple[float | None, str]:
    """dram bytes per launch of the step kernel from the committed ncu capture (one GPU, `ncu --set full`), and what
    that capture was taken from."""
    try:
        with open(os.path.join(REPO, "profiles", "step_kernel_traffic.json")) as f:
            rec = json.load(f)
        return float(rec["dram_bytes_per_launch"]), str(rec.get("source", "profiles/step_kernel_traffic.json"))
    except Exception:
        return None, "no capture committed"


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
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < 1.0:  # a region shorter than nvidia-smi's first sample
            time.sleep(0.05)
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


def base_config(wl_name: str, description: str, particles: int) -> dict:
    """The `config` object both arms print: the workload of the metric and the frame a bench step is."""
    return {"workload": wl_name, "description": description, "particles": particles,
            "steps_per_frame": STEPS_PER_FRAME, "leapfrog_steps_per_bench_step": schedule_steps(STEPS_PER_FRAME),
            "rebins_per_bench_step": 6, "schedule": "reference (kernel_bucket.cuh:181-206)",
            "l2": "state (10M x 20 B x 2 buffers = 400 MB per GPU) is larger than the 126 MB L2; no flush"}


def run_reference(args, rank: int, world: int) -> None:
    """The reference's CPU implementation of the path (Device::CpuThreadPool, all host threads) from oracle/_ref on
    the 10M-particle workload of our arm. Each bench step is a BOUNDED SAMPLE of that workload's frame: `--ref-steps`
    leapfrog steps on the reference schedule (default 18 = one whole re-bin cycle: 1 step, a re-bin, 17 steps) on
    the full scene instead of the frame's 101, so that the run ends within minutes on the box's host cores."""
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
    sample = (f"each bench step = {executed} leapfrog steps + {1 if args.ref_steps > 1 else 0} re-bin of the reference "
              f"schedule (steps_per_frame = {args.ref_steps}) on the full {n}-particle scene, Device::CpuThreadPool, "
              f"{ref.hardware_threads} threads; {args.steps} such steps timed after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": "particle-updates/sec at 10M particles", "value": value,
        "unit": "particle-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": base_config(wl.name, wl.description, n),
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


def slab_parity_check(dist, dev, rank: int, world: int, local_rank: int) -> str:
    """Before anything is timed on more than one GPU: one short frame of the 1M-particle melting liquid (BASELINE.json
    configs[1]; 52 steps, 3 re-bins, migration across every slab boundary) on the `world` slabs, one per rank, and as
    a single slab on rank 0. The slabs' snapshots concatenated in rank order must be byte-identical to the single
    slab's (the check of tests/mp_slab_worker.py, which the driver's one-GPU test box has to skip)."""
    import torch

    from particle_simulator_b200 import slabs, workloads
    from particle_simulator_b200.stepper import Stepper

    wl = workloads.config_1m_liquid()
    wl.frame.metadata["steps_per_frame"] = 52
    n = wl.particles
    uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
    st = Stepper(wl.grid_log2, int(0.75 * n), device=local_rank, slab_rank=rank, slab_count=world, ingest_capacity=n)
    st.comm_init(uid)
    single = Stepper(wl.grid_log2, n, device=local_rank) if rank == 0 else None
    ok = True
    moved = set()
    for s in (st, single):
        if s:
            s.upload(wl.frame)
    for frame in range(2):
        for s in (st, single):
            if s:
                s.run_frame_async()
                s.sync()
        mine = st.download().particles
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = len(mine)
        dist.all_reduce(counts)
        counts = counts.cpu().tolist()
        moved.add(tuple(counts))
        pad = torch.zeros(max(counts) * 20, dtype=torch.uint8, device=dev)
        pad[:len(mine) * 20] = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        if rank == 0:
            got = b"".join(g[:c * 20].cpu().numpy().tobytes() for g, c in zip(gathered, counts))
            ok = ok and got == single.download().particles.tobytes() and sum(counts) == n
    pushed = st.halo_mode == 2
    st.close()
    if single:
        single.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if not int(flag.item()):
        return "FAILED"
    return "bit-identical" + (" (halo pushed by the step kernel" if pushed else " (halo by ncclSend/ncclRecv") + \
        (", particles migrated between slabs)" if len(moved) > 1 else ", no particle changed slab)")


def phase_lines(peak: float) -> list[dict]:
    """BASELINE.json configs[2] is a ramp solid -> liquid -> gas: the step kernel on the same 10M particles and grid in
    the other two regimes (the liquid is the lattice at 1.05 r0 and 150-250 m/s after 3 frames of melting; the gas
    fills the box at 2.4 per cell; both at dt = 10 fs, the step the reference is stable at when hot)."""
    from particle_simulator_b200 import io, workloads
    from particle_simulator_b200.frame import FrameBuffer
    from particle_simulator_b200.stepper import Stepper

    out = []
    for phase in ("liquid", "gas"):
        if phase == "liquid":
            wl = workloads.lattice(3162, 3163, (11, 11), 1.05, 150.0, 250.0, seed=4)
            fb = wl.frame
        else:
            fb = FrameBuffer(10_000_000)
            fb.metadata["box_width"] = fb.metadata["box_height"] = workloads.CELL_WIDTH * 2048
            io.scene_gas(fb, fb.capacity, 2 * workloads.CELL_WIDTH, 3.4e-10, 250.0, 450.0, 0, seed=9)
        fb.metadata["step_dt"] = 10e-15
        fb.metadata["steps_per_frame"] = STEPS_PER_FRAME
        n = fb.count
        with Stepper((11, 11), n) as st:
            st.upload(fb)
            for _ in range(3 if phase == "liquid" else 1):
                st.run_frame_async()
            st.sync()
            st.enable_step_timing(True)
            st.run_frame_async()
            st.sync()
            ms, k = st.step_timing()
            cs = st.cell_start().astype(np.int64)
            stats = st.tile_stats()
        cnt = np.diff(cs).reshape(2048, 2048)
        pad = np.pad(cnt, 1)
        stencil = sum(pad[1 + dy:2049 + dy, 1 + dx:2049 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
        kernel_ms = ms / k
        achieved = ALGO_BYTES_PER_UPDATE * n / (kernel_ms * 1e-3) / 1e9
        out.append({"phase": phase, "particles": n, "kernel_ms": kernel_ms, "achieved": achieved, "frac": achieved / peak,
                    "neighbours_per_particle": float((cnt * (stencil - 1)).sum() / cnt.sum()),
                    "tiles": stats["tiles"], "tiles_staged": stats["tiles_staged"],
                    "threads_live_over_launched": stats["threads_live"] / max(stats["threads_launched"], 1)})
    return out


def run_ours(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist

    from particle_simulator_b200 import slabs, workloads
    from particle_simulator_b200.frame import FrameBuffer, packet_size
    from particle_simulator_b200.stepper import Stepper, balance_rows

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

    slab_parity = None
    if world > 1 and not args.no_slab_parity:
        slab_parity = slab_parity_check(dist, dev, rank, world, local_rank)
        if slab_parity == "FAILED":
            if rank == 0:
                print(json.dumps({"metric": "particle-updates/sec at 10M particles", "value": None, "n_gpus": world,
                                  "slab_parity": "FAILED"}))
            dist.destroy_process_group()
            sys.exit(1)

    # the scene lives in page-locked host memory: it is what the e2e leg uploads every step
    keep = []

    def pinned_frame_storage(count: int) -> np.ndarray:
        t, a = pinned_storage(packet_size(count))
        keep.append(t)
        return a

    stream = torch.cuda.Stream()
    strong = args.scaling == "strong"
    clustered = args.workload == "clustered"
    bounds = None
    if clustered:
        # BASELINE.json configs[4]: 6 droplets (species alternate) + a fast two-species gas, 9.6M particles on 4096^2
        # cells, most droplets in the lower third of the box; rows cut by psim_balance_rows so that every slab holds the
        # same number of particles
        side = args.cluster_side
        wl = workloads.clustered_mixed((12, 12), clusters=6, side=side, gas=side * side * 6 // 9, seed=5)
        if world > 1 and not args.equal_rows:
            bounds = balance_rows(wl.frame, wl.grid_log2[1], world)
        cap = wl.particles if world == 1 else int(1.3 * wl.particles / world) + 65536
        if args.equal_rows:
            cap = wl.particles
        st = Stepper(wl.grid_log2, cap, device=local_rank, slab_rank=rank, slab_count=world, ingest_capacity=wl.particles,
                     bounds=bounds, ghost_capacity=1 << 17 if world > 1 else 0, migrant_capacity=1 << 17 if world > 1 else 0,
                     snapshot_buffers=2, use_graph=world == 1)
    elif strong:
        # strong scaling (BASELINE.json configs[3]): ONE crystal of --total-particles on a fixed 8192 x 4096 grid
        # (box 6.4 x 3.2 um), cut into `world` slabs of 4096 / world cell rows
        rows_log2 = 12 - (world.bit_length() - 1)
        per_slab = args.total_particles // world
        wl = workloads.slab_crystal(rank, world, storage_factory=pinned_frame_storage, per_slab=per_slab,
                                    rows_per_slab_log2=rows_log2, grid_x_log2=13)
        st = Stepper(wl.grid_log2, int(1.05 * per_slab) + 65536, device=local_rank, slab_rank=rank, slab_count=world,
                     ingest_capacity=wl.frame.count, snapshot_buffers=2, use_graph=world == 1)
    elif args.workload == "liquid":
        # BASELINE.json configs[1]: the 1M-particle liquid-density box on 1024^2 cells, one GPU
        if world != 1:
            raise SystemExit("--workload liquid is the single-GPU configuration (BASELINE.json configs[1])")
        wl = workloads.config_1m_liquid(storage=pinned_frame_storage(1000 * 1000))
        st = Stepper(wl.grid_log2, wl.particles, device=local_rank, snapshot_buffers=2, use_graph=True)
    elif world == 1:
        wl = workloads.config_10m_solid(storage=pinned_frame_storage(3162 * 3163))
        st = Stepper(wl.grid_log2, wl.particles, device=local_rank, snapshot_buffers=2, use_graph=True)
    else:
        # weak scaling: one crystal across `world` slabs of 2048 cell rows, ~10M particles per slab; rank r
        # steps slab r, halo rows pushed by the step kernel over NVLink, migrants by NCCL send/recv
        wl = workloads.slab_crystal(rank, world, storage_factory=pinned_frame_storage)
        st = Stepper(wl.grid_log2, int(1.05 * 3162 * 3163), device=local_rank, slab_rank=rank, slab_count=world,
                     ingest_capacity=wl.frame.count, snapshot_buffers=2)
    if world > 1:
        uid = slabs.broadcast_bytes(dist, Stepper.comm_unique_id() if rank == 0 else None, 128, device=dev)
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

    def timed_frames(step_timing: bool):
        """`args.steps` frames chained on the device. Returns (ms, host seconds spent enqueueing, launches, steps)."""
        st.enable_step_timing(step_timing)
        launches0, steps0 = st.kernel_launches, st.steps_executed
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        t_host = time.perf_counter()
        for _ in range(args.steps):
            st.run_frame_async()
        t_host = time.perf_counter() - t_host
        e1.record(stream)
        st.sync()
        barrier()
        return e0.elapsed_time(e1), t_host, st.kernel_launches - launches0, st.steps_executed - steps0

    sampler = ClockSampler(local_rank)
    sampler.start()
    # (1) the number a user gets: no events inside the frames (a single slab replays each frame as one CUDA graph)
    ms, t_host, launches, steps_done = timed_frames(False)
    assert steps_done == executed * args.steps
    # (2) the same frames with a CUDA-event pair around every step-kernel launch: the roofline's kernel time
    ms_ev, _, _, _ = timed_frames(not args.no_step_timing)
    clocks = sampler.stop()
    step_ms, step_launches = st.step_timing()
    st.enable_step_timing(False)
    ms = over_ranks(ms, "max")
    ms_ev = over_ranks(ms_ev, "max")
    value = n * steps_done / (ms * 1e-3)
    held = [int(over_ranks(float(st.particle_count), op)) for op in ("max", "sum")]
    migrants = int(over_ranks(float(st.migrants_sent), "sum"))
    rebins_total = st.rebins_executed

    # ---- end-to-end leg: host frame in, host frame out, every step -------------------------------
    # Pipelined through the public API: every step uploads its scene from page-locked host memory
    # (psim_stage_frame_async + psim_upload_staged) and downloads its result (psim_download_frame_begin / _end); the
    # copies of steps k+1 and k-1 run on their own streams while step k's frame is computed.
    ingest_host_s = []

    def e2e_pipelined(steps: int) -> float:
        barrier()
        t0 = time.perf_counter()
        st.stage_async(wl.frame)
        pending = False
        for k in range(steps):
            t_in = time.perf_counter()
            st.upload_staged()
            ingest_host_s.append(time.perf_counter() - t_in)
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
    ingest_host_s.clear()
    t_e2e = over_ranks(e2e_pipelined(args.steps), "max")
    t_e2e_sync = over_ranks(e2e_synchronous(args.steps), "max")
    assert int(over_ranks(float(out.count), "sum")) == n  # nothing lost, whatever slab holds it now
    e2e_value = n * executed * args.steps / t_e2e
    h2d_local, d2h_local = packet_size(wl.frame.count), packet_size(out.count)
    h2d = int(over_ranks(float(h2d_local), "sum"))
    d2h = int(over_ranks(float(d2h_local), "sum"))
    kernel_ms_max = over_ranks(step_ms / max(step_launches, 1), "max")
    launches = int(over_ranks(float(launches), "sum"))

    # where the end-to-end time goes: the copies alone, all ranks at once (they share the host), and the host time of
    # an ingest (psim_upload_staged returns when the scene is binned: it waits for the frame before it, too)
    def copy_gbs(to_device: bool) -> float:
        devbuf = torch.empty(h2d_local, dtype=torch.uint8, device=dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            a.record()
            for _ in range(3):
                if to_device:
                    devbuf.copy_(host_in, non_blocking=True)
                else:
                    host_out.copy_(devbuf, non_blocking=True)
            b.record()
        torch.cuda.synchronize()
        return 3 * h2d_local / (a.elapsed_time(b) * 1e-3) / 1e9

    host_in = torch.empty(h2d_local, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(h2d_local, dtype=torch.uint8, pin_memory=True)
    h2d_gbs = [over_ranks(copy_gbs(True), op) for op in ("min", "max")]
    d2h_gbs = [over_ranks(copy_gbs(False), op) for op in ("min", "max")]
    ingest_ms = 1e3 * float(np.median(ingest_host_s)) if ingest_host_s else 0.0
    ingest_ms_max = over_ranks(ingest_ms, "max")

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        kernel_ms = step_ms / max(step_launches, 1)
        achieved = ALGO_BYTES_PER_UPDATE * n_local / (kernel_ms * 1e-3) / 1e9  # per GPU (rank 0's slab)
        traffic, traffic_note = recorded_traffic()
        config = base_config(wl.name, wl.description, n)
        if args.workload == "liquid":
            config["l2"] = ("state (1M x (20 B + 16 B of neighbour record) x 2 buffers = 72 MB) FITS the 126 MB L2 and is not flushed "
                            "between frames (a frame is 101 dependent steps): an L2-resident configuration, not the metric's")
        if world > 1 or strong or clustered:
            config["particles_rank0"] = n_local
            config["decomposition"] = ("single slab" if world == 1 else
                                       f"{world} slabs, one per GPU, rows per slab {'cut by psim_balance_rows' if bounds else 'equal'} "
                                       f"(rank 0: {st.slab_info()['rows']}); halo (boundary rows' positions and records) every step: "
                                       f"{halo}; per re-bin: migrants + boundary-row cell counts + fresh ghost rows by ncclSend/ncclRecv")
        line = {
            "metric": "particle-updates/sec at 10M particles", "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "frames": ("each frame replayed as one CUDA graph (PsimConfig.use_graph), no events inside" if world == 1 else
                       "launch by launch, no events inside"),
            "step_kernel": st.tile_stats(),
            "roofline": {"bound": "hbm", "kernel": "step_kernel_c (fused 3x3-cell force + kick + drift)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_note,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * n_local,
                         "kernel_ms": kernel_ms, "kernel_ms_max_over_ranks": kernel_ms_max, "kernel_launches_timed": step_launches,
                         "timed_how": f"a second pass over the same {args.steps} frames with a CUDA-event pair around every "
                                      f"step-kernel launch ({ms_ev / args.steps:.3f} ms per frame)",
                         "kernel_share_of_step": step_ms / ms_ev},
            "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "how": "every step: scene uploaded from page-locked host memory, binned, one frame, snapshot "
                           "downloaded; the copies of steps k+1 / k-1 overlap step k's frame "
                           "(psim_stage_frame_async, psim_upload_staged, psim_download_frame_begin/_end)",
                    "synchronous_ms_per_step": 1e3 * t_e2e_sync / args.steps,
                    "exposed_copy_ms": 1e3 * t_e2e / args.steps - ms / args.steps,
                    "h2d_gbs_min_max_over_ranks": h2d_gbs, "d2h_gbs_min_max_over_ranks": d2h_gbs,
                    "ingest_host_ms": ingest_ms, "ingest_host_ms_max_over_ranks": ingest_ms_max},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": 1e3 * t_host / args.steps,
            "host_numa_binding_rank0": numa,
            "clocks": clocks,
        }
        if slab_parity:
            line["slab_parity"] = slab_parity
        if world > 1 or clustered:
            line["balance"] = {"particles_per_slab_max_over_mean": held[0] / (held[1] / world),
                               "migrants_per_rebin": migrants / max(rebins_total, 1), "rebins": rebins_total}
        if world == 1 and not strong and not clustered and args.workload == "solid":
            st.close()
            if not args.no_phases:
                solid = {"phase": "solid", "particles": n_local, "kernel_ms": kernel_ms, "achieved": achieved,
                         "frac": achieved / peak}
                line["phases"] = [solid] + phase_lines(peak)
            if not args.no_cpu_baseline:  # the reference allocates its own slot arrays on this GPU
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
    ap.add_argument("--ref-steps", type=int, default=18,
                    help="--impl reference: steps_per_frame of one bench step (18 = one re-bin cycle of the schedule)")
    ap.add_argument("--workload", choices=["solid", "liquid", "clustered"], default="solid",
                    help="solid (default): the metric's 10M-particle lattice; liquid: BASELINE.json configs[1], the 1M-particle "
                         "liquid box on one GPU; clustered: BASELINE.json configs[4], droplets + gas "
                         "of two species, rows cut by psim_balance_rows")
    ap.add_argument("--cluster-side", type=int, default=1200, help="--workload clustered: droplets of side x side particles")
    ap.add_argument("--equal-rows", action="store_true", help="--workload clustered: equal rows per slab instead of balanced")
    ap.add_argument("--no-phases", action="store_true", help="skip the liquid / gas step-kernel lines (phases)")
    ap.add_argument("--no-slab-parity", action="store_true", help="N > 1: skip the slabs-vs-single-slab byte comparison")
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
