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
