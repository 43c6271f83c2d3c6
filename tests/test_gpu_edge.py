"""Error behaviour of the per-step cache path where the reference raises IndexError on the host
(model_no_ddp.py:176-179): more forward misses than aux rows, and a sparse id outside its table.
The CUDA path cannot raise from a kernel: it sets a sticky device flag that
``check_device_flags`` turns into the same IndexError, and until then it must stay memory-safe --
the unresolved positions pool a zero row, keep slot -1 and are skipped by the backward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _group(aux, n=1000, d=16, ways=2, csz=8):
    from cdlrm_b200 import model_no_ddp as M
    np.random.seed(0)
    master = M.Embedding_Table_Group(d, np.asarray([n, 50]))
    cg = M.Embedding_Table_Cache_Group(d, np.asarray([n, 50]), csz, aux, ways).to(DEV)
    cg._ensure_ctx(master)
    return master, cg


@pytest.mark.parametrize("one_per_bag", [True, False])
def test_aux_overflow_raises_indexerror_and_stays_memory_safe(one_per_bag):
    B, aux = 64, 5
    master, cg = _group(aux)                      # nothing installed: every lookup misses
    ids = torch.stack([torch.arange(B) * 7 % 1000, torch.arange(B) % 50])
    if one_per_bag:
        lS_o = torch.arange(B).reshape(1, -1).repeat(2, 1)
    else:                                          # general pooling path: 2 ids per bag
        lS_o = (torch.arange(B // 2) * 2).reshape(1, -1).repeat(2, 1)
    opt = torch.optim.SGD(cg.parameters(), lr=0.5)
    before = [e.weight.data.clone() for e in cg.emb_l]
    ly, slots = cg(lS_o, ids, master, 0)
    sum(v.sum() for v in ly).backward()
    opt.step()
    torch.cuda.synchronize()
    for k in range(2):
        sl = slots[k].cpu().numpy()
        base = cg.cache_sizes[k] * cg.num_ways
        assert np.array_equal(sl[:aux], np.arange(base, base + aux))       # misses in batch order (:177)
        assert (sl[aux:] == -1).all()
        out = ly[k].detach().cpu().numpy()
        m = master.emb_l[k].weight.data.numpy()
        if one_per_bag:
            np.testing.assert_array_equal(out[:aux], m[ids[k, :aux].numpy()])
            assert (out[aux:] == 0).all()
        else:
            np.testing.assert_allclose(out[0], m[ids[k, 0]] + m[ids[k, 1]], rtol=1e-6)
            assert (out[3:] == 0).all()
        w = cg.emb_l[k].weight.data
        assert torch.isfinite(w).all()
        assert torch.equal(w[:base], before[k][:base])                     # no stray write into the cache region
    with pytest.raises(IndexError, match="aux"):
        cg.check_device_flags()
    assert cg.check_device_flags() == 0                                    # the flag is cleared by the check


def test_id_outside_table_raises_indexerror():
    B = 32
    master, cg = _group(B)
    ids = torch.stack([torch.arange(B), torch.arange(B) % 50])
    ids[0, 3] = 1000          # == n_rows
    ids[1, 7] = -2
    lS_o = torch.arange(B).reshape(1, -1).repeat(2, 1)
    opt = torch.optim.SGD(cg.parameters(), lr=0.5)
    ly, slots = cg(lS_o, ids, master, 0)
    sum(v.sum() for v in ly).backward()
    opt.step()
    torch.cuda.synchronize()
    assert int(slots[0][3]) == -1 and int(slots[1][7]) == -1
    assert (ly[0][3] == 0).all() and (ly[1][7] == 0).all()
    assert int(cg.last_n_miss[0]) == B and int(cg.last_n_miss[1]) == B
    with pytest.raises(IndexError, match="outside"):
        cg.check_device_flags()


def test_rebind_keeps_planner_tags_and_dirty_bits():
    """Moving / re-binding the cache group after a look-ahead planner exists must not point the planner
    back at the live tags, nor drop the dirty bits of rows touched since the last aggregation."""
    from cdlrm_b200 import cache_manager as C
    B = 32
    master, cg = _group(B)
    planner = C.WindowPlanner(cg, master, 4 * B, rng=C.VictimRngDevice(3, DEV), lookahead_tags=True)
    ids = torch.stack([torch.arange(4 * B) * 3 % 1000, torch.arange(4 * B) % 50]).to(DEV)
    rec = planner.plan(win_ids=ids)                       # plan tags now run ahead of the live tags
    torch.cuda.synchronize()
    assert not all(torch.equal(a, b) for a, b in zip(planner.plan_tags, cg.occupancy_tables))
    cg.dirty_bitmap()[0] = 5
    live_before = [t.clone() for t in cg.occupancy_tables]
    cg.occupancy_tables = [t.clone() for t in cg.occupancy_tables]      # new storage -> key changes -> re-bind
    cg._ensure_ctx(master)
    assert int(cg.dirty_bitmap()[0]) == 5
    rec2 = planner.plan(win_ids=ids)                      # same window again: everything placed now hits
    torch.cuda.synchronize()
    assert sum(rec2.rows) < sum(rec.rows)
    for a, b in zip(cg.occupancy_tables, live_before):    # the live tags were not touched by the plan
        assert torch.equal(a, b)
