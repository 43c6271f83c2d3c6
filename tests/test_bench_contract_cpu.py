"""bench.py's reference arm (the oracle port on the host cores) on the small configuration: stdout must carry
exactly ONE JSON line with the contract's keys; everything else goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "2", "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["unit"] == "samples/s" and r["higher_is_better"] is True
    assert r["value"] > 0 and r["steps"] == 2 and r["n_gpus"] == 1 and r["vs_baseline"] is None
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["sample"]
    assert r["e2e"] == {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and "model" not in r["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
