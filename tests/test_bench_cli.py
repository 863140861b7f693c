"""bench.py's command line, the parts that need no GPU."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, **env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True, cwd=REPO,
                          env=e, timeout=120)


def test_reference_arm_runs_on_rank_0_only():
    """Under torchrun (N > 1) rank 0 alone times the reference; the other ranks leave at once, silently, with 0."""
    proc = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", RANK=1, LOCAL_RANK=1, WORLD_SIZE=2,
               MASTER_ADDR="127.0.0.1", MASTER_PORT=29999)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert proc.stdout.strip() == ""


def test_help_names_the_workloads_of_baseline_json():
    proc = run("--help")
    assert proc.returncode == 0
    for word in ("--gpus", "--steps", "--warmup", "--impl", "liquid", "clustered", "strong"):
        assert word in proc.stdout, word
