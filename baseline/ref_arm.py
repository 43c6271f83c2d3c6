"""bench.py's reference arm: the UNMODIFIED reference (lkp411/cDLRM, copied to baseline/_ref/ by
__graft_entry__.install_reference) timed on the host cores of the GPU box.

What is timed is the reference's own code for the hot path, called the way its ``Run`` loop calls it
(main_no_ddp.py:393-415): ``Prefetcher.process_batch_slice`` + ``CacheEmbeddings`` + the eviction write-back
(cache_manager.py:58-62) per window, and per step ``Embedding_Table_Cache_Group.forward`` -> ``DLRM_Net`` ->
BCE loss -> backward -> ``optimizer_embeds.step()`` / ``optimizer_mlps.step()`` at the FULL batch size.
Nothing of cdlrm_b200 is imported here (no .so is mapped into this process).

Test-side shims only (the reference files are untouched): ``CpuRank('cpu')`` (a str equal to 0, so
``.to(rank)`` stays on the CPU while ``if rank == 0`` fires), ``queue.Queue`` for the eviction fifo, master tables
allocated zero-filled by calloc (lazily committed; their contents do not change the work), the N(0,1) init of the
cache rows skipped while the cache group is constructed (rows are never read before a fill).

Bounded sample (the whole run has to end within minutes): the steps run at the full batch, but a look-ahead
window of ``lookahead`` = 3000 steps (24.6 M ids per table) would take the reference minutes to install, so the
window installed is ``window_steps`` steps long and its measured install time is charged at the TRUE weight of
the configuration, 1/lookahead per step.  A short window holds fewer unique ids than the real one, so this
under-counts the reference's install cost: the figure errs in the reference's favour.
"""
import os
import queue
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return all(os.path.exists(os.path.join(REF, f)) for f in ("main_no_ddp.py", "model_no_ddp.py", "cache_manager.py"))


class CpuRank(str):
    def __eq__(self, o):
        return o == 0 or str.__eq__(self, o)

    __hash__ = str.__hash__


def make_ids(rng, ln_emb, n, dist, zipf_a):
    """Same index distribution as cdlrm_b200/synthetic.py:SyntheticStream (bounded power law / uniform, scrambled)."""
    out = np.empty((len(ln_emb), n), dtype=np.int64)
    for k, nk in enumerate(ln_emb):
        u = rng.random(n)
        if dist == "uniform" or nk == 1:
            r = np.minimum((u * nk).astype(np.int64), nk - 1)
        else:
            e = 1.0 - zipf_a
            r = np.clip((((nk + 1.0) ** e - 1.0) * u + 1.0) ** (1.0 / e) - 1, 0, nk - 1).astype(np.int64)
        out[k] = (r * 2654435761 + 40503 * k) % nk
    return out


def run(wl, ln_emb, steps, warmup, window_steps, dist="zipf", zipf_a=1.05, batch=None, threads=None, log=None):
    """wl: bench.py workload dict (dim, bot, top, batch, cache, ways, lookahead).  Returns a dict with
    ``value`` (samples/s), ``ms_per_step`` and the measured components."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch
    from unittest import mock
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:        # torchrun exports OMP_NUM_THREADS=1: give the OpenMP / BLAS pools all the host cores back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    import cache_manager as C          # the reference's modules (baseline/_ref)
    import main_no_ddp as R
    import model_no_ddp as M
    assert os.path.dirname(os.path.abspath(R.__file__)) == REF, "reference modules must come from baseline/_ref"
    log = log or (lambda m: None)
    d, T, B = int(wl["dim"]), len(ln_emb), int(batch or wl["batch"])
    L_true, Lp = int(wl["lookahead"]), max(1, int(window_steps))
    ln = np.asarray(ln_emb)
    ln_bot = np.fromstring(wl["bot"], dtype=int, sep="-")
    nf = T + 1
    ln_top = np.fromstring(str(nf * (nf - 1) // 2 + int(ln_bot[-1])) + "-" + wl["top"], dtype=int, sep="-")
    np.random.seed(123)
    torch.manual_seed(123)
    # master tables (model_no_ddp.py:21-98) without the 96 GB numpy init: same class, lazily committed storage
    master = M.Embedding_Table_Group.__new__(M.Embedding_Table_Group)
    torch.nn.Module.__init__(master)
    master.emb_l = torch.nn.ModuleList()
    for n in ln_emb:
        EE = torch.nn.EmbeddingBag(int(n), d, mode="sum", sparse=True, _weight=torch.from_numpy(np.zeros((int(n), d), dtype=np.float32)))   # calloc: lazily committed
        EE.weight.requires_grad = False
        master.emb_l.append(EE)
    with mock.patch("torch.nn.init.normal_", lambda t, *a, **k: t):          # skip 2.8 G N(0,1) draws of set-up
        cg = M.Embedding_Table_Cache_Group(d, ln, max_cache_size=int(wl["cache"]), aux_table_size=B,
                                           num_ways=int(wl["ways"]))                              # :346
    dlrm = M.DLRM_Net(ln_bot, ln_top, arch_interaction_op="dot", arch_interaction_itself=False,
                      sigmoid_bot=-1, sigmoid_top=ln_top.size - 2)                                # :351
    loss_fn = torch.nn.BCELoss(reduction="mean")
    opt_m = torch.optim.SGD(dlrm.parameters(), lr=0.8)                                             # :375-376
    opt_e = torch.optim.SGD(cg.parameters(), lr=0.8)
    rank = CpuRank("cpu")
    evq = queue.Queue()
    rng = np.random.default_rng(123)
    lS_o = torch.arange(B).reshape(1, -1).repeat(T, 1)
    n_steps = warmup + steps
    assert Lp >= 1

    def install(win):
        t0 = time.perf_counter()
        rows, uniq, maps = C.Prefetcher.process_batch_slice(win, master)                          # cache_manager.py:27-46
        t1 = time.perf_counter()
        R.CacheEmbeddings(rows, uniq, maps, cg, evq, rank)                                        # main_no_ddp.py:148-209
        t2 = time.perf_counter()
        ev = evq.get()
        for k, (ix, emb) in enumerate(ev):                                                        # cache_manager.py:58-62
            master.emb_l[k].weight.data[ix] = emb
        t3 = time.perf_counter()
        return dict(process_batch_slice_s=t1 - t0, cache_embeddings_s=t2 - t1, writeback_s=t3 - t2,
                    unique_ids=int(sum(u.numel() for u in uniq)), evicted=int(sum(e[0].numel() for e in ev)))

    def one_step(win, b):
        lS_i = win[:, b * B:(b + 1) * B]
        X = torch.from_numpy(np.log1p(rng.integers(0, 101, size=(B, int(ln_bot[0])))).astype(np.float32))
        Y = torch.from_numpy((rng.random((B, 1)) < 0.25).astype(np.float32))
        t0 = time.perf_counter()
        lookups, _idxs = cg(lS_o, lS_i, master, rank)                                             # :403
        Z = dlrm(X, lookups)
        E = loss_fn(Z, Y)
        opt_m.zero_grad()
        opt_e.zero_grad()
        E.backward()
        opt_e.step()
        opt_m.step()
        return time.perf_counter() - t0, float(E.item())

    installs, step_s, losses = [], [], []
    done = 0
    w = 0
    while done < n_steps:
        win = torch.from_numpy(make_ids(rng, ln_emb, Lp * B, dist, zipf_a))
        installs.append(install(win))
        log(f"reference: window {w} installed in {sum(installs[-1][k] for k in ('process_batch_slice_s', 'cache_embeddings_s', 'writeback_s')):.2f} s "
            f"({installs[-1]['unique_ids']} unique ids)")
        for b in range(min(Lp, n_steps - done)):
            t, loss = one_step(win, b)
            if done >= warmup:
                step_s.append(t)
                losses.append(loss)
            done += 1
        w += 1
    inst_s = [i["process_batch_slice_s"] + i["cache_embeddings_s"] + i["writeback_s"] for i in installs]
    step = float(np.mean(step_s))
    per_step = step + min(inst_s) / L_true
    return dict(value=B / per_step, ms_per_step=1000 * per_step, step_ms=1000 * step,
                step_ms_median=1000 * float(np.median(step_s)), install_s=inst_s, installs=installs,
                window_steps=Lp, lookahead=L_true, batch=B, cores=cores, steps=len(step_s), loss_last=losses[-1],
                sample=(f"{len(step_s)} steps of the full {B}-sample batch, all {T} tables at full cardinality, dim {d}, "
                        f"through the reference's own cache_group.forward / DLRM_Net / backward / SGD.step; "
                        f"{len(installs)} window install(s) of {Lp} steps (process_batch_slice + CacheEmbeddings + "
                        f"write-back: {min(inst_s):.2f} s) charged at 1/{L_true} per step (the real window is "
                        f"{L_true} steps: under-counts the reference's install cost); {cores} host threads"))
