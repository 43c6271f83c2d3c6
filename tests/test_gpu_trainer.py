"""Parity of the path bench.py times: ``Trainer`` (look-ahead plan on a side stream by a
background thread, staged install + HBM loser store, lookup on ``forward_stream``, early backward
plan, PDL, flat-bucket dense SGD, fused ``bce_mean``) with the whole step replayed from a CUDA
graph -- against tests/golden/dlrm_trainer.npz, which oracle/gen_golden.py:gen_dlrm_trainer wrote
by running the UNMODIFIED reference loop (main_no_ddp.py:393-415) on the same stream.

Bar: tags after every window bit-exact; miss counts bit-exact; loss curve, final dense
parameters, cache rows and master rows 1e-5 relative (fp32).  Graph replay against the eager
launch of the same step: decisions bit for bit, losses to 1e-6 (the split-K weight gradients and the
cross-warp runs of the sparse update end in floating-point reductions whose order is not fixed, so two
runs of the SAME path already differ in the last bit)."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_trainer(g, use_graph):
    from cdlrm_b200 import main_no_ddp as R
    from cdlrm_b200 import model_no_ddp as M
    cfg = util.golden_cfg(g)
    seed = cfg["seed"]
    ln_emb = np.asarray(cfg["ln_emb"])
    d, B, L, nw = cfg["dim"], cfg["batch"], cfg["lookahead"], cfg["n_windows"]
    T = len(ln_emb)
    args = R.ProcessArgs(["--arch-sparse-feature-size", str(d), "--loss-function", "bce", "--learning-rate",
                          str(cfg["lr_mlp"]), "--lr-embeds", str(cfg["lr_embeds"]), "--mini-batch-size", str(B),
                          "--lookahead", str(L), "--cache-size", str(cfg["cache_size"]), "--num-ways",
                          str(cfg["num_ways"]), "--numpy-rand-seed", str(seed), "--world-size", "1"])
    np.random.seed(seed)
    torch.manual_seed(seed)
    master = M.Embedding_Table_Group(d, ln_emb)                      # numpy RNG order: master first (main :621)
    tr = R.Trainer(args, d, ln_emb, g["ln_bot"], g["ln_top"], master, rank=0, world=1, device=torch.device(DEV))
    # the eager run plans with the reference-shaped eviction lists (their lengths are pinned by the golden), the graph run
    # with the Trainer's default (winners only); caches, master, tags and losses must match the reference either way
    tr.planner.primary_evictions_only = bool(use_graph)
    # flat-bucket mode re-points the parameters: compare by Linear layer order (weights, biases) as the
    # reference's .parameters() yields them
    ref_params = [p for seq in (tr.dlrm.bot_l, tr.dlrm.top_l) for m in seq if isinstance(m, torch.nn.Linear)
                  for p in (m.weight, m.bias)]
    for i, p in enumerate(ref_params):
        assert np.array_equal(p.detach().cpu().numpy(), g[f"mlp_init_{i}"]), "numpy-RNG init order differs"
    ids = util.make_ids(cfg)
    lS_o = torch.arange(B).reshape(1, -1).repeat(T, 1)
    X = torch.from_numpy(g["X"]).to(DEV)
    Y = torch.from_numpy(g["Y"]).to(DEV)
    win = lambda w: torch.from_numpy(ids[:, w * L * B:(w + 1) * L * B]).to(DEV)   # noqa: E731
    losses, tags, n_miss = [], [], []
    tr.submit_window(win(0))
    step = 0
    for w in range(nw):
        rec = tr.install_window()
        if w + 1 < nw:
            tr.submit_window(win(w + 1))            # planned on the side stream while window w trains
        torch.cuda.synchronize()
        tags.append(np.concatenate([t.cpu().numpy().ravel() for t in tr.cache_group.occupancy_tables]))
        # eviction lists: the reference's hold one entry per CLAIMANT of a replaced (set, way) (its length is the golden's
        # evict_len), the Trainer's default ones only the winner of each -- at most one per fill, all flagged primary
        if tr.planner.primary_evictions_only:
            assert all(e <= min(f, ge) for e, f, ge in zip(rec.E, rec.F, g[f"w{w}_evict_len"].tolist()))
            assert all(bool(rec.evict_list(k)[2].all()) for k in range(T))
        else:
            assert rec.E == g[f"w{w}_evict_len"].tolist()
        cur = win(w)
        for b in range(L):
            if use_graph and getattr(tr, "_graph", None) is None and step == 1:
                # capture after one eager step (lazy initialisation), as bench.py does
                tr.capture_graph(X[step], lS_o, cur[:, b * B:(b + 1) * B], Y[step])
            E, _Z = tr.step(X[step], lS_o, cur[:, b * B:(b + 1) * B], Y[step])
            losses.append(E.detach().clone())
            n_miss.append(tr.cache_group.last_n_miss.clone())
            step += 1
    torch.cuda.synchronize()
    tr.finish()          # joins the plan thread, completes the last write-back, raises pending device flags
    if use_graph:
        assert tr._graph is not None and tr.graph_launches > 0
    return dict(losses=np.asarray([float(x) for x in losses], dtype=np.float64),
                tags=tags, n_miss=torch.stack(n_miss).cpu().numpy().astype(np.int64),
                params=[p.detach().cpu().numpy() for p in ref_params],
                weights=[e.weight.data.cpu().numpy() for e in tr.cache_group.emb_l],
                master=[e.weight.data.numpy().copy() for e in master.emb_l])


@pytest.mark.parametrize("use_graph", [True, False])
def test_trainer_matches_reference_golden(use_graph):
    g = util.load_golden("dlrm_trainer.npz")
    cfg = util.golden_cfg(g)
    got = _run_trainer(g, use_graph)
    for w in range(cfg["n_windows"]):
        assert np.array_equal(got["tags"][w], g[f"w{w}_tags"]), f"tags after window {w} differ"
    assert np.array_equal(got["n_miss"], g["n_miss"]), "forward miss counts differ"
    np.testing.assert_allclose(got["losses"], g["losses"], rtol=1e-5)
    for i, p in enumerate(got["params"]):
        util.assert_close_fp32(p, g[f"mlp_final_{i}"], rtol=2e-5, err_msg=f"dense parameter {i}")
    for k in range(len(cfg["ln_emb"])):
        # aux rows are scratch (rewritten by every forward, updates discarded): compare the cache region
        nc = got["weights"][k].shape[0] - cfg["batch"]
        util.assert_close_fp32(got["weights"][k][:nc], g[f"final_weight_{k}"][:nc], err_msg=f"cache rows of table {k}")
        util.assert_close_fp32(got["master"][k], g[f"final_master_{k}"], err_msg=f"master rows of table {k}")


def test_trainer_graph_replay_equals_eager():
    """The captured step and the eager step run the same kernels in the same order on the same
    data: every tag and miss count must agree bit for bit, the loss curve to 1e-6."""
    g = util.load_golden("dlrm_trainer.npz")
    a = _run_trainer(g, True)
    b = _run_trainer(g, False)
    assert np.array_equal(a["n_miss"], b["n_miss"])
    for ta, tb in zip(a["tags"], b["tags"]):
        assert np.array_equal(ta, tb)
    np.testing.assert_allclose(a["losses"], b["losses"], rtol=1e-6, err_msg="graph replay and eager step disagree")
    for pa, pb in zip(a["weights"] + a["params"], b["weights"] + b["params"]):
        np.testing.assert_allclose(pa, pb, rtol=0, atol=1e-6)
