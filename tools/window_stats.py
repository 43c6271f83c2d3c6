#!/usr/bin/env python
"""Developer tool (GPU): plan a few consecutive windows of a bench workload and print, per
window, the planner's counts (unique / hits / dropped / survivor rows / evictions / fills),
the time of each planner phase and the forward-time miss rate of the resulting tags.

  python tools/window_stats.py --workload terabyte --windows 3 [--lookahead N] [--row-cap N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from cdlrm_b200 import cache_manager as C
    from cdlrm_b200 import model_no_ddp as M
    from cdlrm_b200.synthetic import SyntheticStream
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="terabyte")
    ap.add_argument("--windows", type=int, default=3)
    ap.add_argument("--lookahead", type=int, default=0)
    ap.add_argument("--row-cap", type=int, default=40_000_000)
    ap.add_argument("--dist", default="zipf")
    ap.add_argument("--zipf-a", type=float, default=1.05)
    ap.add_argument("--per-table", action="store_true")
    ap.add_argument("--host-rng", action="store_true")
    a = ap.parse_args()
    wl = dict(bench.WORKLOADS[a.workload])
    if a.lookahead:
        wl["lookahead"] = a.lookahead
    ln = bench.table_rows(wl, a.row_cap)
    dev = torch.device("cuda", 0)
    T, L, B, d = len(ln), wl["lookahead"], wl["batch"], wl["dim"]
    cg = M.Embedding_Table_Cache_Group(d, np.asarray(ln), wl["cache"], B, wl["ways"], device=dev)
    cg._ensure_ctx(None)
    pl = C.WindowPlanner(cg, None, L * B, rng=C.VictimRngDevice(123, dev) if not a.host_rng else C.VictimRng(123), lookahead_tags=True)
    st = SyntheticStream(ln, B, dev, dist=a.dist, zipf_a=a.zipf_a, seed=123)
    for w in range(a.windows):
        ids = st.window_ids(w, L)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec = pl.plan(win_ids=ids)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tm = getattr(pl, "last_timing", {})
        # forward-time miss rate against the planned tags, on the first 4 steps of the window
        miss = 0
        n = 4 * B
        for k in range(T):
            tags = pl.plan_tags[k]
            x = ids[k, :n]
            hit = (tags[x % tags.shape[0]] == x[:, None]).any(1)
            miss += int((~hit).sum())
        out = {"window": w, "plan_s": round(dt, 3), "timing": tm, "uniq": sum(rec.uniq), "hits": sum(rec.hits),
               "dropped": sum(rec.dropped), "rows": sum(rec.rows), "E": sum(rec.E), "F": sum(rec.F),
               "fwd_miss_per_step": miss / 4, "fwd_miss_rate": miss / (4 * B * T)}
        print(json.dumps(out))
        if a.per_table:
            for k in range(T):
                print(f"   table {k:2d} n={ln[k]:9d} U={rec.uniq[k]:9d} hit={rec.hits[k]:9d} drop={rec.dropped[k]:8d} "
                      f"R={rec.rows[k]:9d} E={rec.E[k]:9d} F={rec.F[k]:9d}")
        del ids, rec


if __name__ == "__main__":
    main()
