"""GPU parity tests: the CUDA path (through the Python mirror -> ctypes -> C ABI) against the
golden vectors produced by the unmodified reference and against the numpy oracle.

Integer / index / decision results: bit-exact.  Floating point: 1e-5 relative (north_star),
see util.assert_close_fp32."""
import queue

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _mods():
    from cdlrm_b200 import cache_manager as C
    from cdlrm_b200 import main_no_ddp as R
    from cdlrm_b200 import model_no_ddp as M
    return C, R, M


def _run_cuda_trace(cfg, mode):
    """mode 'api': the reference-shaped calls (process_batch_slice -> CacheEmbeddings ->
    eviction data applied to the master -> forward/backward/SGD.step).
    mode 'fast': WindowPlanner on raw window ids with the C++ victim RNG and zero-copy
    evict/fill against the pinned master.  mode 'fast_devrng': the same with the victim stream
    generated on the GPU (cdlrm_rngdev_*)."""
    C, R, M = _mods()
    seed = cfg["seed"]
    np.random.seed(seed)
    torch.manual_seed(seed)
    ln_emb = np.asarray(cfg["ln_emb"])
    d, B, L = cfg["dim"], cfg["batch"], cfg["lookahead"]
    T = len(ln_emb)
    master = M.Embedding_Table_Group(d, ln_emb)
    cg = M.Embedding_Table_Cache_Group(d, ln_emb, max_cache_size=cfg["cache_size"], aux_table_size=B,
                                       num_ways=cfg["num_ways"]).to(DEV)
    opt = torch.optim.SGD(cg.parameters(), lr=cfg["lr_embeds"])
    evq = queue.Queue()
    ids = util.make_ids(cfg)
    grads = util.upstream_grads(cfg)
    rec = util.TraceRecorder()
    lS_o = torch.arange(B).reshape(1, -1).repeat(T, 1)
    torch.manual_seed(seed)
    planner = None
    capped = mode == "staged_capped"      # tiny HBM budget for the loser store: most misses fall back to the host master
    sharded = mode == "staged_sharded"    # loser store cut into 3 per-"rank" shards (all on this device), peer.cu path
    ce = mode == "staged_ce"              # master rows cross PCIe through host threads + cudaMemcpyAsync (hostio.cu)
    primary = mode == "staged_primary"    # eviction lists hold the winner of every replaced (set, way) only (Trainer's setting)
    fillshard = mode == "staged_fillshard"  # fill prefetch cut into 3 per-"rank" shares in 3 staging buffers (peer path)
    if capped or sharded or ce or primary or fillshard:
        mode = "staged"
    if mode in ("fast", "fast_devrng", "staged"):
        cg._ensure_ctx(master)
        rng = C.VictimRng(seed) if mode == "fast" else C.VictimRngDevice(seed, DEV)
        planner = C.WindowPlanner(cg, master, L * B, rng=rng, lookahead_tags=True)
        planner.collect_losers = mode == "staged"
        planner.primary_evictions_only = primary
        if sharded:
            planner.enable_sharded_losers(1, 3, "local")
        if fillshard:
            planner.enable_sharded_fills(1, 3, "local")
        if ce:
            planner.pcie_mode, planner.host_threads = "ce", 3
            planner.CE_CHUNK_BYTES = 64 * 4 * d       # 64-row chunks: many chunk hand-overs even on the tiny traces
    step = 0
    for w in range(cfg["n_windows"]):
        win = torch.from_numpy(ids[:, w * L * B:(w + 1) * L * B])
        if mode == "api":
            rows, uniq, maps = C.Prefetcher.process_batch_slice(win, master)
            R.CacheEmbeddings(rows, uniq, maps, cg, evq, 0)
            ev = evq.get()
            C.Prefetcher.apply_eviction_data(master, ev, cfg.get("avg_wb", False))
            uniq_len = [int(u.numel()) for u in uniq]
            ev_ids = [e[0].numpy() for e in ev]
            ev_rows = [e[1].numpy() for e in ev]
            rng_digest = util.digest(torch.get_rng_state().numpy())
        elif mode == "staged":
            # look-ahead staging: fills and loser rows prefetched into HBM, evictions written back
            # asynchronously, forward misses served from the HBM loser store
            pr = planner.plan(win_ids=win.to(DEV))
            planner.stage(pr)
            planner.install_staged(pr, write_master=True, average_on_writeback=cfg.get("avg_wb", False))
            planner.flush_writeback()         # (copy-engine mode: the evicted rows reach the master here)
            torch.cuda.synchronize()
            if primary:
                assert all(bool(pr.evict_list(k)[2].all()) for k in range(T)) and all(e <= f for e, f in zip(pr.E, pr.F))
            eo = np.concatenate([[0], np.cumsum(pr.E)])
            ev = [(pr.evict_list(k)[0], pr.evict_stage[eo[k]:eo[k] + pr.E[k]]) for k in range(T)]
            for k in range(T):   # losers = window ids that are not cached now, ascending
                u = np.unique(win[k].numpy())
                tags = cg.occupancy_tables[k].cpu().numpy()
                cached = (tags[u % tags.shape[0]] == u[:, None]).any(1)
                assert np.array_equal(pr.loser_list(k).cpu().numpy(), u[~cached][:pr.L[k]])
                assert capped or pr.L[k] == int((~cached).sum())
            assert not capped or sum(pr.L) <= planner.loser_cap_rows
            uniq_len = pr.uniq
            ev_ids = [e[0].cpu().numpy() for e in ev]
            ev_rows = [e[1].cpu().numpy() for e in ev]
        else:
            pr = planner.plan(win_ids=win.to(DEV))
            ev = planner.install(pr, write_master=True, average_on_writeback=cfg.get("avg_wb", False),
                                 collect_evictions=True)
            torch.cuda.synchronize()
            uniq_len = pr.uniq
            ev_ids = [e[0].cpu().numpy() for e in ev]
            ev_rows = [e[1].cpu().numpy() for e in ev]
        if mode != "api":
            rng_digest = None
            for k in range(T):   # the planner's tags and the live tags agree after install
                assert torch.equal(planner.plan_tags[k], cg.occupancy_tables[k])
        rec.window(w, uniq_len, [t.cpu().numpy() for t in cg.occupancy_tables], ev_ids, ev_rows, rng_digest)
        for b in range(L):
            lS_i = win[:, b * B:(b + 1) * B]
            ly, slots = cg(lS_o, lS_i, master, 0)
            G = torch.from_numpy(next(grads)).to(DEV)
            loss = sum((ly[k] * G[k]).sum() for k in range(T))
            opt.zero_grad()
            loss.backward()
            opt.step()
            rec.step(step, torch.stack(slots).cpu().numpy(), cg.last_n_miss.cpu().numpy(),
                     torch.stack([v.detach() for v in ly]).cpu().numpy())
            step += 1
        rec.window_end(w, [e.weight.data.cpu().numpy() for e in cg.emb_l])
    cg.check_device_flags()
    rec.final([e.weight.data.cpu().numpy() for e in cg.emb_l], [e.weight.data.numpy() for e in master.emb_l])
    rec.out["cache_sizes"] = np.asarray(cg.cache_sizes, dtype=np.int64)
    return rec.out


@pytest.mark.parametrize("name", ["trace_tiny.npz", "trace_pressure.npz", "trace_pressure_avgwb.npz",
                                  "trace_cfg0_small.npz"])
@pytest.mark.parametrize("mode", ["api", "fast", "fast_devrng", "staged", "staged_capped", "staged_sharded", "staged_ce",
                                  "staged_primary", "staged_fillshard"])
def test_trace_matches_reference_golden(name, mode, monkeypatch):
    g = util.load_golden(name)
    cfg = util.golden_cfg(g)
    if mode == "staged_capped":           # room for 7 rows in the whole loser store
        monkeypatch.setenv("CDLRM_LOSER_STORE_GB", repr(7 * 4 * cfg["dim"] / 1e9))
    got = _run_cuda_trace(cfg, mode)
    n = util.compare_trace(g, got, check_rng=(mode == "api"), skip_evict_lists=(mode == "staged_primary"))
    if mode == "staged_primary":          # everything else (tags, slots, outputs, final cache and MASTER rows) as the reference
        util.compare_primary_evictions(g, got, cfg["n_windows"])
    assert n > 20


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_interaction_matches_reference_golden(case):
    _, _, M = _mods()
    g = util.load_golden("interact.npz")
    net = M.DLRM_Net.__new__(M.DLRM_Net)
    torch.nn.Module.__init__(net)
    net.arch_interaction_op = "dot"
    net.arch_interaction_itself = bool(g[f"{case}_itself"])
    x = torch.from_numpy(g[f"{case}_x"]).to(DEV).requires_grad_()
    ly = [torch.from_numpy(a).to(DEV).requires_grad_() for a in g[f"{case}_ly"]]
    R = net.interact_features(x, ly)
    util.assert_close_fp32(R.detach().cpu().numpy(), g[f"{case}_R"])
    R.backward(torch.from_numpy(g[f"{case}_dR"]).to(DEV))
    util.assert_close_fp32(x.grad.cpu().numpy(), g[f"{case}_dx"])
    util.assert_close_fp32(torch.stack([t.grad for t in ly]).cpu().numpy(), g[f"{case}_dly"])


@pytest.mark.parametrize("shape", [(2048, 27, 128), (4099, 27, 16), (777, 9, 16), (513, 27, 64), (100, 3, 2),
                                   (8192, 27, 128), (301, 9, 32), (65, 27, 32), (130, 9, 128)])
def test_interaction_matches_oracle_at_size(shape):
    _interaction_at_size(shape)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("shape", [(2048, 27, 128), (513, 27, 64), (301, 9, 32), (130, 9, 128)])
def test_interaction_other_kernel_variants(shape, variant):
    """The mma.sync 3xTF32 kernels (1) and the first CUDA-core forward (2) stay correct (they are
    selectable through cdlrm_interact_set_option and quoted in DESIGN.md as the comparison)."""
    from cdlrm_b200._lib import check, lib
    check(lib.cdlrm_interact_set_option(0, variant))
    try:
        _interaction_at_size(shape)
    finally:
        check(lib.cdlrm_interact_set_option(0, 0))


@pytest.mark.parametrize("fwd", [0, 3, 4])
@pytest.mark.parametrize("shape", [(8192, 27, 128), (2049, 27, 128), (1, 27, 128), (7, 27, 128), (131, 9, 128)])
def test_interaction_forward_kernels_dim128(shape, fwd):
    """Both dim-128 forward kernels against the oracle whichever is the default: 0 = interact_fwd_tr_kernel (a warp
    per sample), 3 = interact_fwd_h_kernel (half a warp per sample; odd batches leave a dead half-warp), 4 = the same
    with persistent, staggered warps (batches above two CTAs per SM; bit-identical to 3)."""
    from cdlrm_b200._lib import check, lib
    check(lib.cdlrm_interact_set_option(2, fwd))
    try:
        _interaction_at_size(shape)
    finally:
        check(lib.cdlrm_interact_set_option(2, -1))


def _interaction_at_size(shape):
    from oracle import oracle as O
    _, _, M = _mods()
    B, nf, d = shape
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, d)).astype(np.float32)
    ly = [rng.standard_normal((B, d)).astype(np.float32) for _ in range(nf - 1)]
    net = M.DLRM_Net.__new__(M.DLRM_Net)
    torch.nn.Module.__init__(net)
    net.arch_interaction_op, net.arch_interaction_itself = "dot", False
    xt = torch.from_numpy(x).to(DEV).requires_grad_()
    lyt = [torch.from_numpy(a).to(DEV).requires_grad_() for a in ly]
    R = net.interact_features(xt, lyt)
    util.assert_close_fp32(R.detach().cpu().numpy(), O.interact_fwd(x, ly))
    dR = rng.standard_normal(tuple(R.shape)).astype(np.float32)
    R.backward(torch.from_numpy(dR).to(DEV))
    dx, dly = O.interact_bwd(x, ly, dR)
    util.assert_close_fp32(xt.grad.cpu().numpy(), dx)
    util.assert_close_fp32(torch.stack([t.grad for t in lyt]).cpu().numpy(), np.stack(dly))


@pytest.mark.parametrize("shape", [(8192, 27, 128), (2049, 27, 128), (1, 27, 128), (131, 9, 128)])
@pytest.mark.parametrize("strided_feats", [False, True])
def test_interaction_pipelined_kernels(shape, strided_feats):
    """The software-pipelined kernels for dim 128 (interact_fwd_pipe_kernel / interact_bwd_pipe_kernel: persistent
    warps / CTAs, bulk async copies into two-stage shared-memory rings).  The backward one is taken when the gradient
    rows are padded to a multiple of 4 floats (what the top MLP's backward hands over).  Both must match the oracle
    and be BIT-identical to the plain kernels (same arithmetic order); odd batch sizes exercise the half-filled last
    item, strided features the per-row copies."""
    from oracle import oracle as O
    from cdlrm_b200._lib import check, lib
    _, _, M = _mods()
    B, nf, d = shape
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, d)).astype(np.float32)
    ly = [rng.standard_normal((B, d)).astype(np.float32) for _ in range(nf - 1)]
    npair = nf * (nf - 1) // 2
    ld = (d + npair + 3) & ~3
    dR = rng.standard_normal((B, d + npair)).astype(np.float32)
    dRp = torch.full((B, ld), float("nan"), device=DEV)[:, :d + npair]      # padded rows, poison in the padding
    dRp.copy_(torch.from_numpy(dR))
    net = M.DLRM_Net.__new__(M.DLRM_Net)
    torch.nn.Module.__init__(net)
    net.arch_interaction_op, net.arch_interaction_itself = "dot", False

    def feats():
        if not strided_feats:
            return [torch.from_numpy(a).to(DEV).requires_grad_() for a in [x] + ly]
        big = torch.zeros(nf, B, d + 4, device=DEV)          # row stride d + 4: not dense
        big[:, :, :d] = torch.from_numpy(np.stack([x] + ly))
        return [big[i, :, :d].detach().requires_grad_() for i in range(nf)]

    grads, outs = {}, {}
    # forward variants: 1 / 2 = two- / one-stage rings, 0 = plain (a warp per sample); 3 = half a warp per sample, 4 = the
    # same persistent, 5 / 6 = the same fed through shared memory by bulk async copies (interact_fwd_hs_kernel)
    for pipe in (1, 0, 2, 3, 4, 5, 6):
        check(lib.cdlrm_interact_set_option(1, min(pipe, 1)))      # backward
        check(lib.cdlrm_interact_set_option(2, pipe))              # forward
        try:
            f = feats()
            R = net.interact_features(f[0], f[1:])
            R.backward(dRp)
            outs[pipe] = R.detach().cpu().numpy()
            grads[pipe] = torch.stack([t.grad for t in f]).cpu().numpy()
        finally:
            check(lib.cdlrm_interact_set_option(1, 1))      # defaults: pipelined backward, the library's forward
            check(lib.cdlrm_interact_set_option(2, -1))
    dx, dly = O.interact_bwd(x, ly, dR)
    util.assert_close_fp32(outs[1], O.interact_fwd(x, ly))
    util.assert_close_fp32(grads[1][0], dx)
    util.assert_close_fp32(grads[1][1:], np.stack(dly))
    assert np.array_equal(outs[1], outs[0]) and np.array_equal(outs[2], outs[0])
    assert np.array_equal(grads[1], grads[0]) and np.array_equal(grads[2], grads[0])
    # the half-warp family: 8-column partials (not bit-identical to the warp-per-sample family), identical among themselves
    util.assert_close_fp32(outs[3], O.interact_fwd(x, ly))
    for v in (4, 5, 6):
        assert np.array_equal(outs[v], outs[3]), f"forward variant {v} differs from variant 3"
        assert np.array_equal(grads[v], grads[0])


@pytest.mark.parametrize("mlp_impl", ["tcgen05", "torch"])
def test_dlrm_tiny_loss_matches_reference_golden(mlp_impl):
    """End to end: cache + interaction kernels + MLPs (tensor-core path and stock PyTorch), BCE
    loss, both optimizers, against the loss curve and final parameters of the reference."""
    C, R, M = _mods()
    g = util.load_golden("dlrm_tiny.npz")
    cfg = util.golden_cfg(g)
    seed = cfg["seed"]
    np.random.seed(seed)
    torch.manual_seed(seed)
    ln_emb = np.asarray(cfg["ln_emb"])
    d, B, L = cfg["dim"], cfg["batch"], cfg["lookahead"]
    T = len(ln_emb)
    master = M.Embedding_Table_Group(d, ln_emb)
    cg = M.Embedding_Table_Cache_Group(d, ln_emb, max_cache_size=cfg["cache_size"], aux_table_size=B,
                                       num_ways=cfg["num_ways"]).to(DEV)
    np.random.seed(seed)
    torch.manual_seed(seed)
    dlrm = M.DLRM_Net(g["ln_bot"], g["ln_top"], arch_interaction_op="dot", arch_interaction_itself=False,
                      sigmoid_bot=-1, sigmoid_top=g["ln_top"].size - 2)
    for i, p in enumerate(dlrm.parameters()):
        assert np.array_equal(p.detach().numpy(), g[f"mlp_init_{i}"])   # same numpy-RNG init order
    dlrm = dlrm.to(DEV)
    dlrm.mlp_impl = mlp_impl
    loss_fn = torch.nn.BCELoss(reduction="mean")
    opt_m = torch.optim.SGD(dlrm.parameters(), lr=0.1)
    opt_e = torch.optim.SGD(cg.parameters(), lr=cfg["lr_embeds"])
    evq = queue.Queue()
    ids = util.make_ids(cfg)
    lS_o = torch.arange(B).reshape(1, -1).repeat(T, 1)
    losses = []
    step = 0
    for w in range(cfg["n_windows"]):
        win = torch.from_numpy(ids[:, w * L * B:(w + 1) * L * B])
        rows, uniq, maps = C.Prefetcher.process_batch_slice(win, master)
        R.CacheEmbeddings(rows, uniq, maps, cg, evq, 0)
        C.Prefetcher.apply_eviction_data(master, evq.get(), False)
        for b in range(L):
            ly, _ = cg(lS_o, win[:, b * B:(b + 1) * B], master, 0)
            Z = dlrm(torch.from_numpy(g["X"][step]).to(DEV), ly)
            E = loss_fn(Z, torch.from_numpy(g["Y"][step]).to(DEV))
            opt_m.zero_grad()
            opt_e.zero_grad()
            E.backward()
            opt_e.step()
            opt_m.step()
            losses.append(E.item())
            step += 1
    np.testing.assert_allclose(np.asarray(losses), g["losses"], rtol=1e-5)
    for i, p in enumerate(dlrm.parameters()):
        util.assert_close_fp32(p.detach().cpu().numpy(), g[f"mlp_final_{i}"], rtol=2e-5)
    for k in range(T):
        util.assert_close_fp32(cg.emb_l[k].weight.data.cpu().numpy(), g[f"final_weight_{k}"])


def test_aggregate_single_rank_matches_reference_golden():
    """W = 1: mean/sum/max of one rank leave the weights unchanged and clear the dirty bits;
    the slot collection equals torch.unique of the idxs (golden rank-0 inputs)."""
    _, R, M = _mods()
    g = util.load_golden("aggregate.npz")
    for op in ("mean", "sum", "max"):
        cg = M.Embedding_Table_Cache_Group(4, np.asarray([50, 7, 300]), max_cache_size=10, aux_table_size=6,
                                           num_ways=2).to(DEV)
        for k, e in enumerate(cg.emb_l):
            e.weight.data.copy_(torch.from_numpy(g[f"{op}_r0_before_{k}"]))
        idx = torch.from_numpy(g[f"{op}_r0_idxs"]).to(DEV)
        R.broadcast_and_aggregate(cg, idx, 0, op)
        torch.cuda.synchronize()
        counts = cg._agg_bufs[2].tolist()
        assert counts == [len(np.unique(g[f"{op}_r0_idxs"][k])) for k in range(3)]
        off = 0
        for k in range(3):
            got = cg._agg_bufs[0][off:off + counts[k]].cpu().numpy()
            assert np.array_equal(got, np.unique(g[f"{op}_r0_idxs"][k]))
            off += cg._cache_rows[k]
            assert np.array_equal(cg.emb_l[k].weight.data.cpu().numpy(), g[f"{op}_r0_before_{k}"])
        assert int(cg.dirty_bitmap().abs().sum()) == 0


def test_device_victim_stream_equals_host_stream():
    """cdlrm_rngdev_* (mt19937 on the GPU + IEEE restatement of glibc's log1p) against
    cdlrm_rng_exponential (std::mt19937 + libm on the host), bit for bit, over calls of awkward
    sizes (state refresh boundaries at multiples of 312 draws) and 20 M draws; and against the
    torch stream recorded in tests/golden/rng.npz."""
    C, _R, _M = _mods()
    g = util.load_golden("rng.npz")
    for seed in (123, 7):    # the torch stream itself, as recorded from the reference's sampler
        r = C.VictimRngDevice(seed, DEV)
        assert np.array_equal(r.exponential(257 * 16).cpu().numpy().reshape(257, 16), g[f"q_{seed}"])
        assert np.array_equal(r.exponential(20).cpu().numpy().reshape(5, 4), g[f"q2_{seed}"])
    seed = 123
    host, dev = C.VictimRng(seed), C.VictimRngDevice(seed, DEV)
    for n in [1, 5, 311, 312, 313, 1, 624, 100_003, 20_000_000, 7]:
        h = host.exponential(n, pin=False)
        d_ = dev.exponential(n).cpu()
        assert torch.equal(h.view(torch.int32), d_.view(torch.int32)), f"draw block of {n} differs"
    assert host.draws == dev.draws
    # hard arguments for log1p: u -> 0 (|x| < 2^-29, < 2^-54, 0) and u -> 1
    raw = []
    for m in [0, 1, 2, (1 << 24) - 1, 1 << 24, (1 << 29) + 12345, (1 << 53) - 1, (1 << 53) - 2, (1 << 52) + 1,
              (1 << 52), 0x000a827999fcef, 0x12bec333018866, 0x15f619980c4337]:
        raw.append(m)
    rng = np.random.default_rng(5)
    raw += [int(x) for x in rng.integers(0, 1 << 53, size=4096)]
    raw += [int(x) >> int(s) for x, s in zip(rng.integers(0, 1 << 53, size=4096), rng.integers(0, 53, size=4096))]
    raw = np.asarray(raw, dtype=np.uint64)
    u = raw.astype(np.float64) * 2.0 ** -53
    want = (-np.log1p(-u)).astype(np.float32)
    got = _device_exp_from_raw(raw)
    assert np.array_equal(want.view(np.int32), got.view(np.int32))


def test_device_victim_stream_jump_ahead_equals_sequential():
    """Chunk-parallel generation of the mt19937 stream (jump-ahead polynomials, csrc/mt_jump_table.h) against the
    sequential single-CTA kernel: same words, same final state, over request sizes that start mid-block, end on and
    off block / chunk boundaries and need up to 8 jump levels (340 M words)."""
    import ctypes
    import time
    from cdlrm_b200._lib import check, lib
    vp = ctypes.c_void_p
    CH = 2555904                                  # words per chunk (MT_JUMP_CHUNK_WORDS), 2 words per draw
    sizes = [1000, 3_000_000, 77, CH, 5, 45_000_000, 700_000, (CH * 3) // 2 + 312 - 1077 // 2, 170_000_000, 11]
    outs = {}
    stream = vp(torch.cuda.current_stream().cuda_stream)
    for mode in ("seq", "par"):
        check(lib.cdlrm_rngdev_set_option(0, -1 if mode == "seq" else 0))
        h = vp()
        check(lib.cdlrm_rngdev_create(ctypes.byref(h), 0, 4242))
        try:
            digests, t_big = [], None
            for n in sizes:
                buf = torch.empty(2 * n, dtype=torch.int32, device=DEV)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                check(lib.cdlrm_rngdev_raw(h, vp(buf.data_ptr()), n, stream))
                torch.cuda.synchronize()
                if n == max(sizes):
                    t_big = time.perf_counter() - t0
                # order-sensitive digest + head / tail words
                w = buf.to(torch.int64) & 0xffffffff
                k = torch.arange(1, w.numel() + 1, device=DEV, dtype=torch.int64)
                digests.append((int(w.sum()), int((w * (k % 65521)).sum() % (1 << 61)), buf[:4].tolist(), buf[-4:].tolist()))
                del buf, w, k
            outs[mode] = (digests, int(lib.cdlrm_rngdev_draws(h)), t_big)
        finally:
            lib.cdlrm_rngdev_destroy(h)
            check(lib.cdlrm_rngdev_set_option(0, 0))
    assert outs["seq"][0] == outs["par"][0]
    assert outs["seq"][1] == outs["par"][1] == sum(sizes)
    print(f"340 M words: sequential {outs['seq'][2] * 1e3:.1f} ms, jump-ahead {outs['par'][2] * 1e3:.1f} ms")
    assert outs["par"][2] < outs["seq"][2]


def _device_exp_from_raw(raw_u64):
    """Runs csrc/expdraw.cuh on given 53-bit integers through the select kernel's twin
    (cdlrm_rngdev_exponential's transform) by planting them as a fake raw stream."""
    import ctypes
    from cdlrm_b200._lib import lib, check
    n = len(raw_u64)
    words = np.empty(2 * n, dtype=np.uint32)
    words[0::2] = (raw_u64 >> np.uint64(32)).astype(np.uint32)
    words[1::2] = (raw_u64 & np.uint64(0xffffffff)).astype(np.uint32)
    d_raw = torch.from_numpy(words.view(np.int32)).to(DEV)
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    check(lib.cdlrm_exp_from_raw(ctypes.c_void_p(d_raw.data_ptr()), ctypes.c_void_p(out.data_ptr()), n,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()
