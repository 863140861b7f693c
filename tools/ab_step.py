#!/usr/bin/env python
"""A/B of step-kernel builds on the bench workload: for every libpsim_<name>.so under build_exp/ (or the names given),
the step-kernel time on the 10M-particle lattice and on the 10M-particle gas, and how far the state after 3 steps is
from the first build's (max position difference in fixed-point units, max velocity difference in m/s).

    python tools/ab_step.py [name ...]          # names of build_exp/libpsim_<name>.so; "product" = the in-tree library
    PSIM_AB_ENV="name:VAR=1" adds environment variables for one name

Each build runs in its own process (PSIM_LIB selects the library).
"""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def child(out_path: str) -> None:
    from particle_simulator_b200 import io, workloads
    from particle_simulator_b200.frame import FrameBuffer
    from particle_simulator_b200.stepper import Stepper

    res = {}
    wl = workloads.config_10m_solid()
    with Stepper(wl.grid_log2, wl.particles, device=0) as st:
        st.upload(wl.frame)
        st.step_async(3)
        st.snapshot_async()
        state = st.download().particles.copy()
        st.step_async(10)
        st.sync()
        st.enable_step_timing(True)
        st.step_async(40)
        st.sync()
        ms, k = st.step_timing()
        res["solid_ms"] = ms / k
        st.enable_step_timing(False)
        # one whole frame, device time
        import time
        st.sync()
        t0 = time.perf_counter()
        st.run_frame_async()
        st.run_frame_async()
        st.sync()
        res["frame_ms"] = (time.perf_counter() - t0) * 500
    try:  # the same frames replayed as CUDA graphs (builds that capture fine-grid frames)
        with Stepper(wl.grid_log2, wl.particles, device=0, use_graph=True) as st:
            st.upload(wl.frame)
            for _ in range(3):
                st.run_frame_async()
            st.sync()
            t0 = time.perf_counter()
            for _ in range(4):
                st.run_frame_async()
            st.sync()
            res["frame_graph_ms"] = (time.perf_counter() - t0) * 250
    except Exception as exc:
        res["frame_graph_ms"] = float("nan")
    np.save(out_path, state)
    n = 10_000_000
    fb = FrameBuffer(n)
    fb.metadata["box_width"] = fb.metadata["box_height"] = workloads.CELL_WIDTH * 2048
    fb.metadata["step_dt"] = 10e-15
    fb.metadata["steps_per_frame"] = 100
    io.scene_gas(fb, n, 2 * workloads.CELL_WIDTH, 3.4e-10, 250.0, 450.0, 0, seed=9)
    with Stepper((11, 11), n) as st:
        st.upload(fb)
        st.run_frame_async()
        st.sync()
        st.enable_step_timing(True)
        st.run_frame_async()
        st.sync()
        ms, k = st.step_timing()
        res["gas_ms"] = ms / k
        stats = st.tile_stats()
        res["gas_tiles"] = stats["tiles"]
    print("RESULT " + json.dumps(res))


def main() -> None:
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    names = sys.argv[1:]
    if not names:
        names = sorted(f[len("libpsim_"):-3] for f in os.listdir(os.path.join(REPO, "build_exp")) if f.startswith("libpsim_"))
    extra = {}
    for item in os.environ.get("PSIM_AB_ENV", "").split():
        name, _, kv = item.partition(":")
        k, _, v = kv.partition("=")
        extra.setdefault(name, {})[k] = v
    base = None
    for name in names:
        env = dict(os.environ)
        libname = name.split("+")[0]
        if libname != "product":
            env["PSIM_LIB"] = os.path.join(REPO, "build_exp", f"libpsim_{libname}.so")
        env.update(extra.get(name, {}))
        out = f"/tmp/ab_{name}.npy"
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", out], env=env, capture_output=True, text=True)
        line = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
        if proc.returncode != 0 or not line:
            print(f"{name:14s} FAILED rc={proc.returncode}\n{proc.stdout[-2000:]}\n{proc.stderr[-2000:]}")
            continue
        res = json.loads(line[0][7:])
        state = np.load(out)
        os.remove(out)
        if base is None:
            base = state
            diff = "(baseline of the comparison)"
        else:
            dx = np.abs((state["x"].astype(np.int64) - base["x"].astype(np.int64) + 2**31) % 2**32 - 2**31).max()
            dy = np.abs((state["y"].astype(np.int64) - base["y"].astype(np.int64) + 2**31) % 2**32 - 2**31).max()
            dv = max(np.abs(state["vx"] - base["vx"]).max(), np.abs(state["vy"] - base["vy"]).max())
            diff = f"vs first after 3 steps: |dx| <= {max(dx, dy)} units, |dv| <= {dv:.3e} m/s"
        print(f"{name:14s} solid {res['solid_ms']:.4f} ms  gas {res['gas_ms']:.4f} ms ({res['gas_tiles']} tiles)  "
              f"frame {res['frame_ms']:.2f} ms (as graphs {res.get('frame_graph_ms', float('nan')):.2f})  {diff}", flush=True)


if __name__ == "__main__":
    main()
