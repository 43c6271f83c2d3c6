"""Shared helpers for the parity tests: synthetic stream, golden loading and
the trace comparison used both for oracle-vs-golden (CPU) and
CUDA-path-vs-oracle/golden (GPU)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_cfg(g):
    return json.loads(str(g["cfg_json"]))


def make_ids(cfg):
    """Synthetic index stream (identical to oracle/gen_golden.py:make_ids)."""
    rng = np.random.default_rng(cfg["data_seed"])
    T = len(cfg["ln_emb"])
    N = cfg["n_windows"] * cfg["lookahead"] * cfg["batch"]
    ids = np.empty((T, N), dtype=np.int64)
    for k, n in enumerate(cfg["ln_emb"]):
        if cfg["dist"] == "uniform":
            ids[k] = rng.integers(0, n, size=N)
        else:
            r = rng.zipf(cfg["zipf_a"], size=N) - 1
            perm_mul = 2654435761 % n if n > 1 else 0
            ids[k] = (r * max(perm_mul, 1) + k) % n
    return ids


def master_init(cfg):
    """Embedding_Table_Group init, model_no_ddp.py:66-74 (numpy global RNG)."""
    np.random.seed(cfg["seed"])
    out = []
    for n in cfg["ln_emb"]:
        W = np.random.uniform(low=-np.sqrt(1 / n), high=np.sqrt(1 / n),
                              size=(n, cfg["dim"])).astype(np.float32)
        out.append(W)
    return out


def upstream_grads(cfg):
    """Generator of the synthetic dL/dV used by gen_golden.run_trace."""
    grng = np.random.default_rng(cfg["data_seed"] + 1)
    T, B, d = len(cfg["ln_emb"]), cfg["batch"], cfg["dim"]
    while True:
        yield grng.standard_normal((T, B, d)).astype(np.float32)


class TraceRecorder:
    """Collects the same keys gen_golden.run_trace writes."""

    def __init__(self):
        self.out = {}

    def window(self, w, uniq_len, tags, evict_ids, evict_rows, rng_digest=None):
        o = self.out
        o[f"w{w}_uniq_len"] = np.asarray(uniq_len, dtype=np.int64)
        o[f"w{w}_tags_digest"] = np.array([digest(t) for t in tags])
        o[f"w{w}_tags"] = np.concatenate([np.asarray(t).ravel() for t in tags])
        o[f"w{w}_evict_len"] = np.asarray([len(e) for e in evict_ids], dtype=np.int64)
        o[f"w{w}_evict_ids"] = np.concatenate([np.asarray(e, dtype=np.int64) for e in evict_ids])
        o[f"w{w}_evict_rows"] = np.concatenate([np.asarray(r, dtype=np.float32) for r in evict_rows], axis=0)
        o[f"w{w}_evict_rows_sum"] = np.asarray([np.asarray(r, dtype=np.float64).sum() for r in evict_rows])
        if rng_digest is not None:
            o[f"w{w}_rng_digest"] = np.array(rng_digest)

    def step(self, s, slots, n_miss, outs):
        o = self.out
        sl = np.asarray(slots, dtype=np.int32)
        o[f"s{s}_n_miss"] = np.asarray(n_miss, dtype=np.int64)
        o[f"s{s}_slots_digest"] = np.array(digest(sl))
        o[f"s{s}_slots"] = sl
        o[f"s{s}_out"] = np.asarray(outs, dtype=np.float32)
        o[f"s{s}_out_sum"] = np.asarray([np.asarray(v, dtype=np.float64).sum() for v in outs])

    def window_end(self, w, weights):
        self.out[f"w{w}_weight_sum"] = np.asarray([np.asarray(x, dtype=np.float64).sum() for x in weights])

    def final(self, weights, master):
        for k, (wt, m) in enumerate(zip(weights, master)):
            self.out[f"final_weight_{k}"] = np.asarray(wt)
            self.out[f"final_weight_s_{k}"] = np.asarray(wt)[::97]
            self.out[f"final_master_{k}"] = np.asarray(m)
            self.out[f"final_master_s_{k}"] = np.asarray(m)[::97]
            self.out[f"final_master_sum_{k}"] = np.asarray(np.asarray(m, dtype=np.float64).sum())


def assert_close_fp32(v, ref, rtol=1e-5, err_msg=""):
    """north_star tolerance: 1e-5 relative (fp32).  Entries that are the result of
    cancellation (w - lr*g near 0) cannot meet an element-wise relative bound, so
    the absolute floor is 1e-5 x the tensor's own scale (max |ref|)."""
    ref = np.asarray(ref)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    np.testing.assert_allclose(np.asarray(v), ref, rtol=rtol, atol=rtol * max(scale, 1e-30), err_msg=err_msg)


def compare_primary_evictions(golden, got, n_windows, rtol=1e-5):
    """Eviction lists emitted with one entry per replaced (set, way) (WindowPlanner.primary_evictions_only) against the
    reference's lists, which hold one entry per claimant (main_no_ddp.py:190-199): per window and table the same SET of
    evicted ids, each once, with the row the reference evicted for it."""
    for w in range(n_windows):
        g_len, m_len = golden[f"w{w}_evict_len"], got[f"w{w}_evict_len"]
        g_ids, m_ids = golden[f"w{w}_evict_ids"], got[f"w{w}_evict_ids"]
        g_rows, m_rows = golden.get(f"w{w}_evict_rows"), got[f"w{w}_evict_rows"]    # (large fixtures keep only the sums)
        go = np.concatenate([[0], np.cumsum(g_len)])
        mo = np.concatenate([[0], np.cumsum(m_len)])
        for k in range(len(g_len)):
            gi, mi = g_ids[go[k]:go[k + 1]], m_ids[mo[k]:mo[k + 1]]
            u, first = np.unique(gi, return_index=True)
            assert len(mi) == len(u) and np.array_equal(np.sort(mi), u), f"window {w} table {k}: evicted id set differs"
            if len(u) and g_rows is not None:
                order = np.argsort(mi)
                assert_close_fp32(m_rows[mo[k]:mo[k + 1]][order], g_rows[go[k]:go[k + 1]][first], rtol=rtol,
                                  err_msg=f"w{w} table {k} evicted rows")


def compare_trace(golden, got, rtol=1e-5, check_rng=True, skip_evict_lists=False):
    """Bit-exact for every integer/decision key, ``rtol`` (north_star: 1e-5
    relative, fp32) for floats.  Only keys present in the golden are checked."""
    checked = 0
    for key, ref in golden.items():
        if key in ("cfg_json", "master_init_digest") or key.startswith("master_init_"):
            continue
        if skip_evict_lists and "_evict_" in key:
            continue
        if key.endswith("_rng_digest") and not check_rng:
            continue
        assert key in got, f"missing key {key}"
        val = got[key]
        if ref.dtype.kind in ("U", "S"):
            assert np.array_equal(ref, val), f"{key}: digest mismatch"
        elif ref.dtype.kind in ("i", "u", "b"):
            assert np.array_equal(ref, np.asarray(val)), f"{key}: integer mismatch"
        else:
            v = np.asarray(val)
            assert v.shape == ref.shape, f"{key}: shape {v.shape} vs {ref.shape}"
            if key.endswith("_sum"):
                # sums of many fp32 values: scale the tolerance by the magnitude summed
                np.testing.assert_allclose(v, ref, rtol=1e-4, atol=1e-2, err_msg=key)
            else:
                assert_close_fp32(v, ref, rtol=rtol, err_msg=key)
        checked += 1
    assert checked > 0
    return checked


def run_oracle_trace(cfg):
    """The oracle (oracle/oracle.py) driven exactly like gen_golden.run_trace
    drives the reference."""
    from oracle import oracle as O

    master = master_init(cfg)
    T, B, L, d = len(cfg["ln_emb"]), cfg["batch"], cfg["lookahead"], cfg["dim"]
    cache = O.OracleCache(d, cfg["ln_emb"], cfg["cache_size"], B, cfg["num_ways"])
    gen = O.TorchCpuGenerator(cfg["seed"])
    ids = make_ids(cfg)
    grads = upstream_grads(cfg)
    rec = TraceRecorder()
    offsets = np.arange(B, dtype=np.int64)
    step = 0
    for w in range(cfg["n_windows"]):
        win = ids[:, w * L * B:(w + 1) * L * B]
        ev, plans, uniq = O.install_window(cache, master, win, gen, cfg.get("avg_wb", False))
        rec.window(w, [len(u) for u in uniq], cache.tags, [e[0] for e in ev], [e[1] for e in ev])
        for b in range(L):
            lS_i = win[:, b * B:(b + 1) * B]
            ly, slots, n_miss = O.forward(cache, [offsets] * T, lS_i, master)
            rec.step(step, slots, n_miss, ly)
            G = next(grads)
            for k in range(T):
                O.backward_sgd_table(cache.weight[k], slots[k], offsets, G[k], cfg["lr_embeds"])
            step += 1
        rec.window_end(w, cache.weight)
    rec.final(cache.weight, master)
    rec.out["cache_sizes"] = np.asarray(cache.cache_sizes, dtype=np.int64)
    return rec.out
