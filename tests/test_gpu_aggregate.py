"""Table aggregation (main_no_ddp.py:250-292) for W = 2 on ONE GPU, against the reference's own
2-rank result (tests/golden/aggregate.npz, written by oracle/gen_golden.py:gen_aggregate, which ran
the unmodified ``broadcast_and_aggregate`` on two gloo ranks).

Two cache-group replicas live on the same device; two host threads call the product's
``broadcast_and_aggregate`` -- mark -> bitmap all-gather + OR -> collect -> pack -> all-reduce ->
unpack, every kernel of the multi-rank path -- with the NCCL transport replaced by an in-process
one (same stream, host barriers).  The multi-GPU NCCL run of the same function is
tools/mgpu_check.py (tests/test_multi_gpu.py, needs 2 GPUs)."""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Shared:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world, timeout=120)
        self.slot = [None] * world


class ThreadComm:
    """all_gather_into / all_reduce between replicas that share a device and a stream: enqueue order
    is execution order, so host barriers between the phases are all the synchronisation needed."""

    def __init__(self, shared, rank):
        self.sh, self.rank = shared, rank

    @property
    def world(self):
        return self.sh.world

    def all_gather_into(self, out, inp):
        sh = self.sh
        sh.slot[self.rank] = inp
        sh.barrier.wait()
        n = inp.numel()
        for r in range(sh.world):
            out[r * n:(r + 1) * n].copy_(sh.slot[r])
        sh.barrier.wait()

    def all_reduce(self, buf, op):
        sh = self.sh
        sh.slot[self.rank] = buf
        sh.barrier.wait()
        assert all(b.shape == buf.shape for b in sh.slot), "ranks packed different numbers of rows"
        st = torch.stack(list(sh.slot))
        red = st.max(0).values if op == "max" else st[0] + st[1] if sh.world == 2 else st.sum(0)
        sh.barrier.wait()
        buf.copy_(red)
        sh.barrier.wait()


def _groups(g, op, W=2):
    from cdlrm_b200 import model_no_ddp as M
    cgs = []
    for r in range(W):
        cg = M.Embedding_Table_Cache_Group(4, np.asarray([50, 7, 300]), max_cache_size=10, aux_table_size=6,
                                           num_ways=2).to(DEV)
        for k, e in enumerate(cg.emb_l):
            e.weight.data.copy_(torch.from_numpy(g[f"{op}_r{r}_before_{k}"]))
        cg._ensure_ctx(None)
        cgs.append(cg)
    return cgs


@pytest.mark.parametrize("op", ["mean", "sum", "max"])
@pytest.mark.parametrize("via", ["idxs", "dirty"])
def test_two_rank_aggregate_matches_reference_golden(op, via):
    """via 'idxs': the reference's call (explicit int32 slot tensors); via 'dirty': the slots were
    marked beforehand (what the fused backward does) and ``cache_group_idxs`` is None."""
    from cdlrm_b200 import main_no_ddp as R
    from cdlrm_b200._lib import check, lib
    import ctypes
    g = util.load_golden("aggregate.npz")
    W = 2
    cgs = _groups(g, op, W)
    idxs = [torch.from_numpy(g[f"{op}_r{r}_idxs"]).to(DEV) for r in range(W)]
    if via == "dirty":
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for r in range(W):
            check(lib.cdlrm_agg_mark(cgs[r]._ctx, ctypes.c_void_p(idxs[r].data_ptr()), idxs[r].stride(0),
                                     idxs[r].shape[1], s))
    torch.cuda.synchronize()
    sh = _Shared(W)

    def work(r):
        torch.cuda.set_device(0)
        R.broadcast_and_aggregate(cgs[r], idxs[r] if via == "idxs" else None, r, op, comm=ThreadComm(sh, r))
        return True

    with ThreadPoolExecutor(W) as ex:
        assert all(f.result(timeout=300) for f in [ex.submit(work, r) for r in range(W)])
    torch.cuda.synchronize()
    union = [np.unique(np.concatenate([g[f"{op}_r{r}_idxs"][k] for r in range(W)])) for k in range(3)]
    for r in range(W):
        counts = cgs[r]._agg_bufs[2].tolist()
        assert counts == [len(u) for u in union], "slot union differs from torch.unique of the gathered idxs"
        off = 0
        for k in range(3):
            got = cgs[r]._agg_bufs[0][off:off + counts[k]].cpu().numpy()
            assert np.array_equal(got, union[k])             # ascending, as torch.unique(sorted=True) (:270)
            off += cgs[r]._cache_rows[k]
            w = cgs[r].emb_l[k].weight.data.cpu().numpy()
            ref = g[f"{op}_r{r}_after_{k}"]
            if op == "max":
                assert np.array_equal(w, ref), f"rank {r} table {k}"
            else:
                np.testing.assert_allclose(w, ref, rtol=1e-6, atol=1e-7, err_msg=f"rank {r} table {k}")
        assert int(cgs[r].dirty_bitmap().abs().sum()) == 0, "dirty bits must be cleared by the aggregation"
    # both replicas hold identical rows on the union afterwards
    for k in range(3):
        u = torch.from_numpy(union[k]).to(DEV).long()
        assert torch.equal(cgs[0].emb_l[k].weight.data[u], cgs[1].emb_l[k].weight.data[u])


def test_three_replica_aggregate_matches_oracle():
    """W = 3, larger tables, random dirty sets: against oracle.aggregate (numpy restatement of :250-292)."""
    from cdlrm_b200 import main_no_ddp as R
    from cdlrm_b200 import model_no_ddp as M
    from oracle import oracle as O
    W, d = 3, 16
    ln_emb = np.asarray([5000, 37, 90000])
    rng = np.random.default_rng(4)
    cgs, before, idxs = [], [], []
    for r in range(W):
        cg = M.Embedding_Table_Cache_Group(d, ln_emb, max_cache_size=211, aux_table_size=32, num_ways=4).to(DEV)
        for e in cg.emb_l:
            e.weight.data.copy_(torch.from_numpy(rng.standard_normal(tuple(e.weight.shape)).astype(np.float32)))
        cg._ensure_ctx(None)
        cgs.append(cg)
        before.append([e.weight.data.cpu().numpy().copy() for e in cg.emb_l])
        rows = min(e.weight.shape[0] for e in cg.emb_l)
        idxs.append(rng.integers(0, rows, size=(3, 700)).astype(np.int32))
    sh = _Shared(W)

    def work(r):
        torch.cuda.set_device(0)
        R.broadcast_and_aggregate(cgs[r], torch.from_numpy(idxs[r]).to(DEV), r, "mean", comm=ThreadComm(sh, r))
        return True

    with ThreadPoolExecutor(W) as ex:
        assert all(f.result(timeout=300) for f in [ex.submit(work, r) for r in range(W)])
    torch.cuda.synchronize()
    O.aggregate(before, idxs, "mean")
    for r in range(W):
        for k in range(3):
            np.testing.assert_allclose(cgs[r].emb_l[k].weight.data.cpu().numpy(), before[r][k], rtol=1e-6, atol=1e-7)
