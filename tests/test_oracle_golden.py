"""Pins oracle/oracle.py (the numpy restatement) against golden vectors produced
by the UNMODIFIED reference (oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

import util
from oracle import oracle as O


def test_geometry_matches_reference():
    with open(os.path.join(util.GOLDEN, "geometry.json")) as f:
        g = json.load(f)
    for s, want in g["find_next_prime"].items():
        assert O.find_next_prime(int(s)) == want, s
    for n, want in g["isPrime"].items():
        assert O.is_prime_ref(int(n)) == want, n
    # the quirky values
    assert O.find_next_prime(10000) == 10006 and O.find_next_prime(150000) == 150001
    assert O.find_next_prime(300000) == 300002


def test_rng_restatement_matches_torch_stream():
    g = util.load_golden("rng.npz")
    for seed in (123, 7):
        gen = O.TorchCpuGenerator(seed)
        q = gen.exponential_f32(257 * 16).reshape(257, 16)
        assert np.array_equal(q, g[f"q_{seed}"])
        q2 = gen.exponential_f32(20).reshape(5, 4)       # split-invariant stream
        assert np.array_equal(q2, g[f"q2_{seed}"])


@pytest.mark.parametrize("ways", [4, 16, 5])
def test_way_sampler_matches_categorical(ways):
    g = util.load_golden("rng.npz")
    avail = g[f"avail_{ways}"]
    q = O.TorchCpuGenerator(99).exponential_f32(avail.size).reshape(avail.shape)
    assert np.array_equal(O.sample_ways(avail, q), g[f"sample_{ways}"])


@pytest.mark.parametrize("name", ["trace_tiny.npz", "trace_pressure.npz", "trace_pressure_avgwb.npz",
                                  "trace_cfg0_small.npz"])
def test_trace_matches_reference(name):
    g = util.load_golden(name)
    cfg = util.golden_cfg(g)
    m = util.master_init(cfg)
    assert [util.digest(x) for x in m] == list(g["master_init_digest"])
    got = util.run_oracle_trace(cfg)
    n = util.compare_trace(g, got, check_rng=False)
    assert n > 20


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_interaction_matches_reference(case):
    g = util.load_golden("interact.npz")
    x, ly = g[f"{case}_x"], list(g[f"{case}_ly"])
    itself = bool(g[f"{case}_itself"])
    R = O.interact_fwd(x, ly, itself)
    np.testing.assert_allclose(R, g[f"{case}_R"], rtol=1e-5, atol=1e-5)
    dx, dly = O.interact_bwd(x, ly, g[f"{case}_dR"], itself)
    np.testing.assert_allclose(dx, g[f"{case}_dx"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.stack(dly), g[f"{case}_dly"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("op", ["mean", "sum", "max"])
def test_aggregate_matches_reference(op):
    g = util.load_golden("aggregate.npz")
    weights = [[g[f"{op}_r{r}_before_{k}"].copy() for k in range(3)] for r in range(2)]
    idxs = [g[f"{op}_r{r}_idxs"] for r in range(2)]
    O.aggregate(weights, idxs, op)
    for r in range(2):
        for k in range(3):
            np.testing.assert_allclose(weights[r][k], g[f"{op}_r{r}_after_{k}"], rtol=1e-6, atol=1e-7)


def test_fast_variants_equal_reference_variants():
    """The vectorised functions timed by bench.py's cpu_baseline give the same results."""
    rng = np.random.default_rng(2)
    w1 = rng.standard_normal((300, 8)).astype(np.float32)
    w2 = w1.copy()
    slots = rng.integers(0, 300, size=256).astype(np.int32)
    dV = rng.standard_normal((256, 8)).astype(np.float32)
    O.backward_sgd_table(w1, slots, np.arange(256), dV, 0.3)
    O.backward_sgd_table_fast(w2, slots, dV, 0.3)
    np.testing.assert_allclose(w1, w2, rtol=1e-5, atol=1e-6)
    x = rng.standard_normal((7, 16)).astype(np.float32)
    ly = [rng.standard_normal((7, 16)).astype(np.float32) for _ in range(5)]
    R1 = O.interact_fwd(x, ly)
    R2, T = O.interact_fwd_fast(x, ly)
    np.testing.assert_allclose(R1, R2, rtol=1e-5, atol=1e-5)
    dR = rng.standard_normal(R1.shape).astype(np.float32)
    dx, dly = O.interact_bwd(x, ly, dR)
    dT = O.interact_bwd_fast(T, dR)
    np.testing.assert_allclose(dT[:, 0], dx, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dT[:, 1:].transpose(1, 0, 2), np.stack(dly), rtol=1e-5, atol=1e-5)
    cache = O.OracleCache(8, [1000, 40], 64, 32, 4)
    master = [rng.standard_normal((n, 8)).astype(np.float32) for n in (1000, 40)]
    ids = np.stack([rng.integers(0, n, size=32) for n in (1000, 40)])
    O.install_window(cache, master, ids, O.TorchCpuGenerator(1))
    for k in range(2):
        a, sa, ma = O.forward_table(cache, k, np.arange(32), ids[k], master[k])
        b, sb, mb = O.forward_table_fast(cache, k, ids[k], master[k])
        assert np.array_equal(sa, sb) and ma == mb and np.array_equal(a, b)


def test_trainer_golden_decisions_match_oracle():
    """dlrm_trainer.npz (the reference's Run loop body end to end): the tag evolution and the
    forward miss counts depend on the index stream and the victim generator only, so the oracle
    must reproduce them without the MLPs."""
    g = util.load_golden("dlrm_trainer.npz")
    cfg = util.golden_cfg(g)
    master = util.master_init(cfg)
    T, B, L, d = len(cfg["ln_emb"]), cfg["batch"], cfg["lookahead"], cfg["dim"]
    cache = O.OracleCache(d, cfg["ln_emb"], cfg["cache_size"], B, cfg["num_ways"])
    gen = O.TorchCpuGenerator(cfg["seed"])
    ids = util.make_ids(cfg)
    off = np.arange(B, dtype=np.int64)
    n_miss = []
    for w in range(cfg["n_windows"]):
        win = ids[:, w * L * B:(w + 1) * L * B]
        ev, _plans, _uniq = O.install_window(cache, master, win, gen)
        assert np.array_equal(np.concatenate([t.ravel() for t in cache.tags]), g[f"w{w}_tags"])
        assert [len(e[0]) for e in ev] == g[f"w{w}_evict_len"].tolist()
        for b in range(L):
            _ly, _slots, nm = O.forward(cache, [off] * T, win[:, b * B:(b + 1) * B], master)
            n_miss.append(nm)
    assert np.array_equal(np.asarray(n_miss, dtype=np.int64), g["n_miss"])


@pytest.mark.parametrize("name", ["trace_pressure.npz", "trace_pressure_avgwb.npz", "trace_tiny.npz"])
def test_reference_eviction_lists_repeat_one_row_per_replaced_slot(name):
    """What WindowPlanner.primary_evictions_only (cdlrm_plan_set_primary_evictions) relies on, checked on the REFERENCE's
    own output: when several ids of a window claim the same (set, way), the reference's eviction list repeats the old
    tag once per claimant (main_no_ddp.py:190-199) and every repeat carries the same cache row -- so writing each
    evicted id back once (cache_manager.py:48-64, plain or averaged) leaves the master exactly as the reference does."""
    g = util.load_golden(name)
    cfg = util.golden_cfg(g)
    repeats = 0
    for w in range(cfg["n_windows"]):
        ln, ids, rows = g[f"w{w}_evict_len"], g[f"w{w}_evict_ids"], g[f"w{w}_evict_rows"]
        o = np.concatenate([[0], np.cumsum(ln)])
        for k in range(len(ln)):
            gi, gr = ids[o[k]:o[k + 1]], rows[o[k]:o[k + 1]]
            u, inv = np.unique(gi, return_inverse=True)
            repeats += len(gi) - len(u)
            first = np.full(len(u), -1, dtype=np.int64)
            for i in range(len(gi) - 1, -1, -1):
                first[inv[i]] = i
            assert np.array_equal(gr, gr[first[inv]]), f"window {w} table {k}: repeats of an evicted id carry different rows"
    if "pressure" in name:
        assert repeats > 50             # the undersized cache does produce duplicate claims
