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
    import math
    B, W = 10, 4
    lb = math.ceil(B / W)
    ids = torch.arange(3 * B).reshape(3, B)
    got = torch.cat([ids[:, r * lb:(r + 1) * lb] for r in range(W)], dim=1)
    assert torch.equal(got, ids)
