"""bench.py's reference arm on the small configuration (the reference's own code from baseline/_ref when
__graft_entry__.install_reference has copied it -- the build container and the GPU box --, else the numpy oracle
port): stdout must carry exactly ONE JSON line with the contract's keys; everything else goes to stderr; the arm
must not map the product's CUDA library into its process."""
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
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "main_no_ddp.py"))
    assert r["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["sample"]
    if have_ref:
        assert r["cpu_baseline"]["batch"] == 128 and r["cpu_baseline"]["lookahead"] == 100     # the full batch of the config
    assert r["e2e"] == {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and "model" not in r["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_reference_arm_does_not_load_the_product_library():
    """The CPU arm times the reference: nothing of cdlrm_b200 (in particular libcdlrm_b200.so) may be imported."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'small', '--steps', '1', "
            "'--warmup', '0']; runpy.run_path(%r, run_name='__main__'); "
            "bad = [m for m in sys.modules if m.startswith('cdlrm_b200.') and m != 'cdlrm_b200.synthetic']; "
            "maps = open('/proc/self/maps').read(); "
            "assert not bad and 'libcdlrm_b200' not in maps, (bad, 'libcdlrm_b200' in maps)") % os.path.join(ROOT, "bench.py")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
