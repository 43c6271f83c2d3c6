"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/cdlrm_b200.h declares, the host-side functions (geometry, victim RNG) match the
reference's golden vectors, and the Python mirror keeps the reference's CLI."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import util


def _header_symbols():
    src = open(os.path.join(util.ROOT, "include", "cdlrm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cdlrm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cdlrm_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 35
    for name in syms:
        assert hasattr(_lib.lib, name), f"libcdlrm_b200.so does not export {name}"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert _lib.lib.cdlrm_abi_version() == 1


def test_geometry_matches_reference_golden():
    from cdlrm_b200._lib import lib
    g = json.load(open(os.path.join(util.GOLDEN, "geometry.json")))
    for s, want in g["find_next_prime"].items():
        got = lib.cdlrm_find_next_prime(int(s))
        assert (None if got < 0 else got) == want, s
    for n, want in g["isPrime"].items():
        assert bool(lib.cdlrm_is_prime_ref(int(n))) == want, n


def test_victim_rng_matches_torch_stream_golden():
    from cdlrm_b200.cache_manager import VictimRng
    g = util.load_golden("rng.npz")
    for seed in (123, 7):
        for threads in (1, 4):
            r = VictimRng(seed, threads=threads)
            q = r.exponential(257 * 16, pin=False).numpy().reshape(257, 16)
            assert np.array_equal(q, g[f"q_{seed}"])
            q2 = r.exponential(20, pin=False).numpy().reshape(5, 4)
            assert np.array_equal(q2, g[f"q2_{seed}"])
            assert r.draws == 257 * 16 + 20


def test_victim_rng_threaded_equals_serial_on_large_draw():
    from cdlrm_b200.cache_manager import VictimRng
    a = VictimRng(5, threads=1).exponential(3_000_000, pin=False)
    b = VictimRng(5, threads=6).exponential(3_000_000, pin=False)
    assert np.array_equal(a.numpy(), b.numpy())
    import torch
    torch.manual_seed(5)
    assert np.array_equal(a.numpy(), torch.empty(3_000_000).exponential_(1).numpy())


def test_cli_flags_and_defaults_match_reference():
    from cdlrm_b200.main_no_ddp import ProcessArgs
    g = json.load(open(os.path.join(util.GOLDEN, "geometry.json")))["cli_defaults"]
    mine = vars(ProcessArgs([]))
    for k, v in g.items():
        assert k in mine, f"flag {k} missing"
        assert mine[k] == v, (k, mine[k], v)
    a = ProcessArgs("--cache-size 150000 --num-ways 16 --lookahead 3000 --cache-workers 4 --table-agg-freq 100 "
                    "--batch-fifo-size 8 --mini-batch-size 8192 --large-batch".split())
    assert (a.cache_size, a.num_ways, a.lookahead, a.cache_workers, a.table_agg_freq, a.batch_fifo_size) == \
        (150000, 16, 3000, 4, 100, 8)


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cdlrm_b200 import _lib
    from cdlrm_b200.model_no_ddp import Embedding_Table_Cache_Group, Embedding_Table_Group
    np.random.seed(0)
    master = Embedding_Table_Group(8, np.asarray([100, 20]))
    cg = Embedding_Table_Cache_Group(8, np.asarray([100, 20]), 16, 4, 2)
    with pytest.raises(_lib.CdlrmError):
        cg(torch.arange(4).repeat(2, 1), torch.zeros(2, 4, dtype=torch.long), master, 0)
    h = ctypes.c_void_p()
    rc = _lib.lib.cdlrm_ctx_create(ctypes.byref(h), 0, 2, 8, 2, 4, _lib.i64_array([100, 20]), 16)
    assert rc != 0 and b"" != _lib.lib.cdlrm_last_error()
    # the whole-window prefetch / write-back pipelines need the copy engine: an error without a device, and argument
    # errors (no staging chunks, negative counts) before anything is touched
    M = np.zeros((10, 8), np.float32)
    ids = np.arange(4, dtype=np.int64)
    chunk = np.zeros((4, 8), np.float32)
    vp = ctypes.c_void_p
    args = (0, 1, _lib.ptr_array([M.ctypes.data]), _lib.i64_array([10]), 8, _lib.ptr_array([ids.ctypes.data]),
            _lib.i64_array([4]), _lib.ptr_array([chunk.ctypes.data]))
    assert _lib.lib.cdlrm_host_prefetch_rows(*args, vp(chunk.ctypes.data), vp(chunk.ctypes.data), 4, 1, None) != 0
    assert _lib.lib.cdlrm_host_prefetch_rows(*args, None, None, 4, 1, None) != 0
    assert np.array_equal(chunk, np.zeros((4, 8), np.float32))
    assert _lib.lib.cdlrm_host_writeback_rows(0, 1, _lib.ptr_array([M.ctypes.data]), _lib.i64_array([10]), 8,
                                              _lib.ptr_array([ids.ctypes.data]), None, _lib.i64_array([4]),
                                              _lib.ptr_array([chunk.ctypes.data]), vp(chunk.ctypes.data),
                                              vp(chunk.ctypes.data), 4, 0, 1, None) != 0
    assert not M.any()


def test_module_attribute_surface_matches_reference():
    from cdlrm_b200.model_no_ddp import DLRM_Net, Embedding_Table_Cache_Group, Embedding_Table_Group, isPrime
    np.random.seed(1)
    cg = Embedding_Table_Cache_Group(4, np.asarray([50, 7, 300]), max_cache_size=10, aux_table_size=6, num_ways=2)
    assert cg.max_cache_size == 10 and cg.cache_sizes == [10, 7, 10]   # isPrime(10) is True in the reference
    assert [tuple(e.weight.shape) for e in cg.emb_l] == [(26, 4), (20, 4), (26, 4)]
    assert all(t.shape == (s, 2) and int(t.min()) == -1 for t, s in zip(cg.occupancy_tables, cg.cache_sizes))
    assert len(list(cg.parameters())) == 3 and cg.victim_cache_entries == [None] * 3
    assert int(cg.compute_set_indices(0, __import__("torch").tensor([23]))[0]) == 3
    m = Embedding_Table_Group(4, np.asarray([50, 7]))
    assert m.emb_l[0].weight.shape == (50, 4) and not m.emb_l[0].weight.requires_grad
    assert float(m.emb_l[1].weight.abs().max()) <= np.sqrt(1 / 7) + 1e-6
    d = DLRM_Net(np.asarray([13, 8, 4]), np.asarray([10, 8, 1]), "dot", sigmoid_top=1)
    assert len(d.bot_l) == 4 and len(d.top_l) == 4 and hasattr(d, "interact_features")
    assert isPrime(10006) and isPrime(9) and not isPrime(15)   # the reference quirks


def test_host_gather_scatter_rows_match_numpy():
    """hostio.cu: the host half of the copy-engine prefetch / write-back (cache_manager.py:34-43 `weight[unique_idxs]`,
    :58-62 `weight[idxs] = rows` and its averaged variant) against numpy, single- and multi-threaded."""
    from cdlrm_b200._lib import check, lib
    vp = ctypes.c_void_p
    rng = np.random.default_rng(0)
    n_rows, d = 50_000, 16
    M = rng.standard_normal((n_rows, d)).astype(np.float32)
    ids = rng.integers(0, n_rows, 30_000).astype(np.int64)
    for threads in (1, 5):
        dst = np.empty((ids.size, d), np.float32)
        check(lib.cdlrm_host_gather_rows(vp(M.ctypes.data), n_rows, d, vp(ids.ctypes.data), ids.size, vp(dst.ctypes.data), threads))
        assert np.array_equal(dst, M[ids])
        u = np.unique(ids)
        rows = rng.standard_normal((u.size, d)).astype(np.float32)
        prim = (rng.random(u.size) < 0.6).astype(np.uint8)
        for average in (0, 1):
            got = M.copy()
            check(lib.cdlrm_host_scatter_rows(vp(got.ctypes.data), n_rows, d, vp(u.ctypes.data), vp(prim.ctypes.data), u.size,
                                              vp(rows.ctypes.data), average, threads))
            want = M.copy()
            sel = prim == 1
            want[u[sel]] = (want[u[sel]] + rows[sel]) / 2 if average else rows[sel]
            assert np.array_equal(got, want)
    bad = np.asarray([0, n_rows], dtype=np.int64)      # an id outside its table is an error, not a wild access
    dst = np.zeros((2, d), np.float32)
    assert lib.cdlrm_host_gather_rows(vp(M.ctypes.data), n_rows, d, vp(bad.ctypes.data), 2, vp(dst.ctypes.data), 1) != 0


def test_flat_bucket_layout_for_the_early_allreduce():
    """DLRM_Net.flatten_parameters: bottom-MLP weights, top-MLP weights, then the biases; the top MLP's weights are the
    contiguous range [flat_top_weight_off, flat_weight_elems) that Trainer all-reduces as soon as its backward is
    enqueued (biases are never reduced: main_no_ddp.py:234-247)."""
    import torch
    from cdlrm_b200.model_no_ddp import DLRM_Net
    np.random.seed(1)
    net = DLRM_Net(np.asarray([13, 7, 5]), np.asarray([9, 6, 1]), "cat", sigmoid_top=1)
    flat_p, flat_g = net.flatten_parameters()
    pad = lambda n: (n + 3) & ~3
    bot = pad(13 * 7) + pad(7 * 5)
    top = pad(9 * 6) + pad(6 * 1)
    assert net.flat_top_weight_off == bot and net.flat_weight_elems == bot + top
    assert flat_p.numel() == bot + top + pad(7) + pad(5) + pad(6) + pad(1)
    lin = [m for seq in (net.bot_l, net.top_l) for m in seq if isinstance(m, torch.nn.Linear)]
    assert lin[2].weight.data_ptr() == flat_p.data_ptr() + 4 * bot              # first top-MLP weight
    assert lin[0].bias.data_ptr() == flat_p.data_ptr() + 4 * (bot + top)        # first bias right behind the weights
    assert all(m.weight.grad.data_ptr() - flat_g.data_ptr() == m.weight.data_ptr() - flat_p.data_ptr() for m in lin)
