"""CPU: `bench.py --impl reference` (the oracle port on the host cores, persistent workers) prints one JSON line with the
contract's keys and a positive throughput; non-zero ranks exit without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--horizon", "4"], capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["steps_requested"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BalanceCartPole+boundary_projection")
    assert d["worker_errors"]["count"] == 0
    # the arm states what it actually ran (the label said 4096 environments, the run had 128)
    cfg = d["config"]
    assert cfg["envs_per_worker"] == 64 and cfg["num_envs_cpu"] == 64 * cfg["cpu_workers"] == 64 * d["cpu_baseline"]["cores"]
    assert d["gates"]["nonfinite_env_steps"] == 0 and d["gates"]["env_steps"] == cfg["num_envs_cpu"] * 4 * d["steps"]
    assert d["cpu_baseline"]["ms_per_env_step_per_core"] > 0


def test_reference_arm_other_ranks_do_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
