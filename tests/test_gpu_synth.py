"""Device-side synthetic id stream (cdlrm_synth_ids) and the chunked window scan (cdlrm_plan_mark_ids):
slices of the counter-based stream agree with the whole window, the distribution is what the host loaders
draw from, and a window marked chunk by chunk plans exactly like the same window handed over as one tensor."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dist", ["zipf", "uniform"])
def test_stream_slices_agree_and_ids_are_in_range(dist):
    from cdlrm_b200.synthetic import SyntheticStream
    ln = [39884406, 3, 1, 155, 585935]
    Bg, L = 96, 7
    st = SyntheticStream(ln, Bg, DEV, dist=dist, zipf_a=1.05, seed=5)
    g = st.window_ids(2, L)                                            # steps 14..20, all samples
    assert g.shape == (len(ln), L * Bg) and g.dtype == torch.int64
    for k, n in enumerate(ln):
        assert int(g[k].min()) >= 0 and int(g[k].max()) < n
    v = g.view(len(ln), L, Bg)
    assert torch.equal(st.ids(2 * L + 3, 2), v[:, 3:5].reshape(len(ln), -1))               # step sub-range
    assert torch.equal(st.ids(2 * L, L, b0=32, nb=32), v[:, :, 32:64].reshape(len(ln), -1))  # a rank's slice
    assert torch.equal(SyntheticStream(ln, Bg, DEV, dist=dist, seed=5).window_ids(2, L), g)  # a pure function of the seed
    assert not torch.equal(SyntheticStream(ln, Bg, DEV, dist=dist, seed=6).window_ids(2, L), g)
    big = SyntheticStream([1_000_000], 1 << 16, DEV, dist=dist, zipf_a=1.05, seed=1).ids(0, 8)[0]
    uniq = int(torch.unique(big).numel())
    if dist == "uniform":
        assert uniq > 0.38 * big.numel()         # 524288 draws from 1 M ids: ~41 % distinct
    else:
        top = torch.bincount(big, minlength=1_000_000).max().item() / big.numel()
        assert 0.05 < top < 0.2 and uniq < 0.3 * big.numel()           # a power law: one id takes ~9 % of the draws


def test_chunked_scan_plans_like_the_whole_window():
    from cdlrm_b200 import cache_manager as C
    from cdlrm_b200 import model_no_ddp as M
    from cdlrm_b200.synthetic import SyntheticStream
    ln = np.asarray([20000, 37, 6000, 400000])
    d, Bg, L = 16, 256, 9
    st = SyntheticStream(ln, Bg, DEV, dist="zipf", zipf_a=1.05, seed=3)
    recs, tags = [], []
    for chunked in (False, True):
        np.random.seed(0)
        master = M.Embedding_Table_Group(d, ln)
        cg = M.Embedding_Table_Cache_Group(d, ln, 300, Bg, 8).to(DEV)
        cg._ensure_ctx(master)
        pl = C.WindowPlanner(cg, master, L * Bg, rng=C.VictimRngDevice(7, DEV), lookahead_tags=True)
        pl.collect_losers = True
        out = []
        for w in range(3):
            if not chunked:
                rec = pl.plan(win_ids=st.window_ids(w, L))
            else:
                for s0 in range(0, L, 4):        # chunks of 4, 4 and 1 steps
                    pl.mark_ids(st.ids(w * L + s0, min(4, L - s0)))
                rec = pl.plan(marked=L * Bg)
            torch.cuda.synchronize()
            lists = []
            for k in range(len(ln)):
                lists += [t.cpu().numpy() for t in rec.fill_list(k)] + [t.cpu().numpy() for t in rec.evict_list(k)]
                lists.append(rec.loser_list(k).cpu().numpy())
            out.append([rec.uniq, rec.hits, rec.dropped, rec.rows, rec.E, rec.F, rec.L] + lists)
        recs.append(out)
        tags.append([t.cpu().numpy() for t in pl.plan_tags])
    for a, b in zip(*recs):
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    for a, b in zip(*tags):
        assert np.array_equal(a, b)
    assert sum(recs[0][2][4]) > 0, "the stream was meant to cause evictions by the third window"
