#!/usr/bin/env python
"""Which ingredient breaks the FIRST CUDA-graph capture of a Trainer step in a fresh process?
Each variant runs in its own subprocess: python tools/capture_probe.py [variant]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = {
    "base_1eager": dict(eager=1),
    "eager3": dict(eager=3),
    "no_pdl": dict(eager=1, env={"CDLRM_PDL": "0"}),
    "mlp_torch": dict(eager=1, env={"CDLRM_MLP": "torch"}),
    "no_flat": dict(eager=1, env={"CDLRM_FLAT_MLP": "0"}),
    "no_fused_loss": dict(eager=1, env={"CDLRM_FUSED_LOSS": "0"}),
    "no_fwd_stream": dict(eager=1, fwd_stream=False),
    "no_early_plan": dict(eager=1, early_plan=False),
    "prio_stream": dict(eager=1, prio=True),
    "plan_before": dict(eager=1, submit_next=False),
    "pool_streams": dict(eager=1, env={"CDLRM_POOL_STREAMS": "1"}),
    "keepE": dict(eager=1, keepE=True),
    "sync_tags": dict(eager=1, sync_tags=True),
    "clones": dict(eager=1, clones=True),
    "all3": dict(eager=1, keepE=True, sync_tags=True, clones=True),
    "as_test": dict(as_test=True),
    "pytest_plain": dict(pytest=["-q", "-x"]),
    "pytest_nocapture": dict(pytest=["-q", "-x", "-s"]),
    "pytest_nofault": dict(pytest=["-q", "-x", "-p", "no:faulthandler"]),
}


def child(name):
    v = VARIANTS[name]
    if v.get("as_test"):
        import test_gpu_trainer as tt
        tt.test_trainer_matches_reference_golden(True)
        print(f"{name}: OK")
        return
    if v.get("pytest"):
        import pytest
        rc = pytest.main(v["pytest"] + [os.path.join(ROOT, "tests", "test_gpu_trainer.py") + "::test_trainer_matches_reference_golden"])
        print(f"{name}: OK" if rc == 0 else f"{name}: pytest Error rc={rc}")
        sys.exit(int(rc))
    import numpy as np
    import torch
    import util
    from cdlrm_b200 import main_no_ddp as R
    from cdlrm_b200 import model_no_ddp as M
    g = util.load_golden("dlrm_trainer.npz")
    cfg = util.golden_cfg(g)
    ln_emb = np.asarray(cfg["ln_emb"])
    d, B, L = cfg["dim"], cfg["batch"], cfg["lookahead"]
    T = len(ln_emb)
    args = R.ProcessArgs(["--arch-sparse-feature-size", str(d), "--loss-function", "bce", "--mini-batch-size", str(B),
                          "--lookahead", str(L), "--cache-size", str(cfg["cache_size"]), "--num-ways",
                          str(cfg["num_ways"]), "--world-size", "1"])
    np.random.seed(1)
    master = M.Embedding_Table_Group(d, ln_emb)
    dev = torch.device("cuda:0")
    if v.get("prio"):
        torch.cuda.set_stream(torch.cuda.Stream(dev, priority=-1))
    tr = R.Trainer(args, d, ln_emb, g["ln_bot"], g["ln_top"], master, rank=0, world=1, device=dev)
    if v.get("fwd_stream") is False:
        tr.cache_group.forward_stream = None
    if v.get("early_plan") is False:
        tr.cache_group.early_plan = False
    ids = util.make_ids(cfg)
    lS_o = torch.arange(B).reshape(1, -1).repeat(T, 1)
    X = torch.from_numpy(g["X"]).to(dev)
    Y = torch.from_numpy(g["Y"]).to(dev)
    win = lambda w: torch.from_numpy(ids[:, w * L * B:(w + 1) * L * B]).to(dev)   # noqa: E731
    tr.submit_window(win(0))
    tr.install_window()
    if v.get("submit_next", True):
        tr.submit_window(win(1))
    if v.get("sync_tags"):
        torch.cuda.synchronize()
        _tags = [t.cpu().numpy().ravel() for t in tr.cache_group.occupancy_tables]
    cur = win(0)
    keep = []
    for s in range(v["eager"]):
        out = tr.step(X[s], lS_o, cur[:, s * B:(s + 1) * B], Y[s])
        if v.get("keepE"):
            keep.append(out)
        if v.get("clones"):
            keep.append((out[0].detach().clone(), tr.cache_group.last_n_miss.clone()))
        del out
    s = v["eager"]
    tr.capture_graph(X[s], lS_o, cur[:, s * B:(s + 1) * B], Y[s])
    E, _ = tr.step(X[s], lS_o, cur[:, s * B:(s + 1) * B], Y[s])
    torch.cuda.synchronize()
    print(f"{name}: OK loss {float(E):.5f}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for name, v in VARIANTS.items():
            env = dict(os.environ, **v.get("env", {}))
            r = subprocess.run([sys.executable, __file__, name], env=env, capture_output=True, text=True, timeout=300)
            tail = [ln for ln in (r.stdout + r.stderr).splitlines() if "Error" in ln or ": OK" in ln]
            print(f"{name:16s} rc={r.returncode}  {tail[-1] if tail else ''}")
            if r.returncode != 0:
                tb = [ln for ln in r.stderr.splitlines() if ln.strip().startswith("File") and "cdlrm_b200" in ln]
                print("      " + " | ".join(x.strip() for x in tb[-3:]))
