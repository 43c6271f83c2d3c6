"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: the flat-bucket MLP
gradient all-reduce (main_no_ddp.py:234-247 semantics: weights averaged, biases untouched)
and the batch slicing contract (:388-391)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cdlrm_b200.main_no_ddp import aggregate_gradients, wait_wrap
    from cdlrm_b200.model_no_ddp import DLRM_Net
    np.random.seed(3)
    net = DLRM_Net(np.asarray([5, 4, 2]), np.asarray([3, 4, 1]), "cat", sigmoid_top=1)
    g = torch.Generator().manual_seed(10 + rank)
    for p in net.parameters():
        p.grad = torch.randn(p.shape, generator=g)
    before = [p.grad.clone() for p in net.parameters()]
    wait_wrap(aggregate_gradients(net))
    ret.put((rank, [b.numpy() for b in before], [p.grad.numpy().copy() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_gradients_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29733, ret)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted([ret.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    [p.join() for p in ps]
    (_, b0, a0), (_, b1, a1) = got
    for i in range(len(b0)):
        if b0[i].ndim == 2:       # Linear weight: averaged over ranks
            np.testing.assert_allclose(a0[i], (b0[i] / 2 + b1[i] / 2), rtol=1e-6)
            np.testing.assert_allclose(a1[i], a0[i], rtol=0, atol=0)
        else:                      # bias: the reference never reduces it
            np.testing.assert_array_equal(a0[i], b0[i])
            np.testing.assert_array_equal(a1[i], b1[i])


def test_batch_slicing_contract():
    """main_no_ddp.slice_batch (:388-391), the function Run cuts every batch with: the ranks' slices tile the
    global batch in rank order; offsets are the first local_batch columns."""
    import math

    from cdlrm_b200.main_no_ddp import slice_batch
    B, W, Tn = 10, 4, 3
    lb = math.ceil(B / W)
    X = torch.arange(B * 13, dtype=torch.float32).reshape(B, 13)
    ids = torch.arange(Tn * B).reshape(Tn, B)
    lS_o = torch.arange(B).reshape(1, -1).repeat(Tn, 1)
    Y = torch.arange(B, dtype=torch.float32).reshape(B, 1)
    parts = [slice_batch(r, lb, X, lS_o, ids, Y) for r in range(W)]
    assert torch.equal(torch.cat([p[0] for p in parts]), X)
    assert torch.equal(torch.cat([p[2] for p in parts], dim=1), ids)
    assert torch.equal(torch.cat([p[3] for p in parts]), Y)
    for r, (x, o, i, y) in enumerate(parts):
        n = max(0, min(lb, B - r * lb))
        assert x.shape[0] == i.shape[1] == y.shape[0] == n
        assert torch.equal(o, lS_o[:, :lb])                      # :390: offsets are NOT re-based (P = 1)


def test_prefetcher_window_slicing_contract():
    """Prefetcher.windows (cache_manager.py:75,85-110): FIFO entry w covers exactly the training steps
    [w*lookahead, (w+1)*lookahead) of the epoch (main_no_ddp.py:393), the last one may be short, and every
    epoch starts a new window; cache_ld is a twin of train_ld (same ids in the same order)."""
    import argparse

    from cdlrm_b200.cache_manager import Prefetcher
    from cdlrm_b200.synthetic import make_synthetic_data_and_loaders
    ln_emb = [50, 7, 3000]
    args = argparse.Namespace(lookahead=4, nepochs=2, mini_batch_size=8, num_batches=10, data_size=1,
                              numpy_rand_seed=5, test_mini_batch_size=-1, cache_workers=1, main_start_core=0,
                              average_on_writeback=False, eviction_fifo_timeout=1)
    train_ld, test_ld, cache_ld = make_synthetic_data_and_loaders(args, ln_emb, 13)
    assert len(train_ld) == len(cache_ld) == 10
    cm = Prefetcher(args, None, None, None, None, cache_ld)
    assert cm.fifo_payload == "tuples"                           # the reference's contract is the default
    wins = list(cm.windows())
    assert [w.shape[1] for w in wins] == [32, 32, 16] * 2
    steps = [b[2] for b in train_ld]
    for e in range(2):
        for w in range(3):
            want = torch.cat(steps[w * 4:(w + 1) * 4], dim=1)
            assert torch.equal(wins[e * 3 + w], want)
    for k, n in enumerate(ln_emb):
        assert int(wins[0][k].min()) >= 0 and int(wins[0][k].max()) < n
    x, o, i, t = next(iter(test_ld))
    assert x.shape == (8, 13) and i.shape == (3, 8) and t.shape == (8, 1) and torch.equal(o[0], torch.arange(8))


def test_writeback_shares_partition_the_eviction_list():
    """Trainer.install_window at N > 1: rank r writes back wb_share_range(E, r, W) of every table's eviction list
    (cache_manager.py:48-64 is done by the one eviction manager in the reference): the shares must tile [0, E)."""
    from cdlrm_b200.cache_manager import wb_share_range
    for E in (0, 1, 7, 8, 1000, 123457):
        for W in (1, 2, 3, 8):
            parts = [wb_share_range(E, r, W) for r in range(W)]
            assert parts[0][0] == 0 and sum(n for _lo, n in parts) == E
            for (lo, n), (lo2, _n2) in zip(parts, parts[1:]):
                assert lo + n == lo2
            assert max(n for _lo, n in parts) - min(n for _lo, n in parts) <= 1
