"""The JSON line bench.py prints (the driver's contract), checked on its quickest workload: BASELINE.json configs[1],
the 1M-particle liquid box (`--workload liquid`), two timed frames. Named to run after the parity tests."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_has_the_contract_keys():
    proc = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--workload", "liquid", "--steps", "2", "--warmup", "3"],
                          capture_output=True, text=True, cwd=REPO, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, proc.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["unit"] == "particle-updates/s" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "1M-liquid" and "l2" in d["config"]
    # value = particles x leapfrog steps per frame / time per frame
    n, steps = d["config"]["particles"], d["config"]["leapfrog_steps_per_bench_step"]
    assert d["value"] == pytest.approx(n * steps / (d["ms_per_step"] * 1e-3), rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.0 < r["frac"] < 1.0
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(40 * n / (r["kernel_ms"] * 1e-3) / 1e9, rel=1e-6)  # 40 algorithmic bytes per update
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0
    assert e["h2d_bytes_per_step"] >= 20 * n and e["d2h_bytes_per_step"] >= 20 * n  # a 20-byte record each way, every step
    assert d["gpu_launches"] >= 2 * steps  # our own kernels ran inside the timed region
    assert "sm_mhz" in d["clocks"] and "sm_max_mhz" in d["clocks"] and isinstance(d["clocks"]["reasons"], list)
    assert d["clocks"]["sm_mhz"] is None or d["clocks"]["sm_mhz"] > 0  # None only if nvidia-smi gave no sample at all
