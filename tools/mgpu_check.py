#!/usr/bin/env python
"""Multi-GPU consistency check of the data-parallel cache path (run under torchrun, one rank
per GPU; used by tests/test_multi_gpu.py and by hand through gpurun --gpus N):

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py

Trains a small DLRM for a few windows through ``Trainer`` (look-ahead plan on every rank,
staged install, loser store, periodic table aggregation over NCCL) and asserts the
invariants of the replicated-cache design:
  * tags are bit-identical on all ranks after every window boundary (same deterministic plan);
  * right after a table aggregation the cached rows (non-aux region) are bit-identical on all
    ranks (touched rows were averaged, untouched rows were identical before);
  * the averaged rows equal the mean of the per-rank rows taken just before the aggregation
    (checked against a torch all_gather of the rows, fp32 1e-6);
  * rank 0's write-back reaches the shared master: after the last boundary every evicted id's
    master row equals the (averaged) row that was evicted;
  * Linear weights stay identical across ranks (grads all-reduced every step); loss is finite.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from cdlrm_b200.main_no_ddp import ProcessArgs, Trainer, broadcast_and_aggregate
    from cdlrm_b200.model_no_ddp import Embedding_Table_Group
    from cdlrm_b200.synthetic import SyntheticStream
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ln_emb = [5000, 37, 90000, 1200]
    d, lb, L, ways, csz, agg = 16, 96, 5, 4, 211, 3
    Bg = lb * world
    T = len(ln_emb)
    args = ProcessArgs(["--arch-sparse-feature-size", str(d), "--arch-mlp-bot", f"13-32-{d}", "--arch-mlp-top", "32-1",
                        "--loss-function", "bce", "--learning-rate", "0.1", "--lr-embeds", "0.3",
                        "--mini-batch-size", str(Bg), "--lookahead", str(L), "--cache-size", str(csz), "--num-ways",
                        str(ways), "--table-agg-freq", str(agg), "--world-size", str(world)])
    ln_bot = np.asarray([13, 32, d])
    nf = T + 1
    ln_top = np.asarray([nf * (nf - 1) // 2 + d, 32, 1])
    prefix = f"/dev/shm/cdlrm_mgpu_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        master = Embedding_Table_Group(d, np.asarray(ln_emb), init=f"shm:{prefix}:create")
    dist.barrier()
    if rank != 0:
        master = Embedding_Table_Group(d, np.asarray(ln_emb), init=f"shm:{prefix}:attach")
    tr = Trainer(args, d, np.asarray(ln_emb), ln_bot, ln_top, master, rank=rank, world=world, device=dev)
    dist.barrier()
    if rank == 0:
        for k in range(T):
            os.unlink(f"{prefix}_{k}.bin")
    sg = SyntheticStream(ln_emb, Bg, dev, dist="zipf", zipf_a=1.05, seed=11)
    sl = SyntheticStream(ln_emb, lb, dev, dist="zipf", zipf_a=1.05, seed=100 + rank)
    lS_o = torch.arange(lb).reshape(1, -1).repeat(T, 1)
    cg = tr.cache_group
    n_windows = 4

    def gathered(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    use_marker = os.environ.get("MGPU_MARKER", "0") == "1"      # chunked, rank-sharded window scan (bench.py's path)

    def submit(w, g_):
        if not use_marker:
            return tr.submit_window(g_)

        def mark(planner):
            r_, w_ = planner.scan_shard
            per = (L + w_ - 1) // w_
            for s0 in range(r_ * per, min(L, (r_ + 1) * per), 2):
                ns = min(2, min(L, (r_ + 1) * per) - s0)
                planner.mark_ids(sg.ids(w * L + s0, ns, stream=planner.stream))
            return L * Bg
        tr.submit_window(mark, own_ids=g_.view(T, L, world, lb)[:, :, rank].reshape(T, L * lb))

    g = sg.window_ids(0, L)
    submit(0, g)
    checked_losers = 0
    j = 0
    losses = []
    for w in range(n_windows):
        rec = tr.install_window()
        torch.cuda.synchronize()
        for k in range(T):                       # (1) tags identical on every rank
            for other in gathered(cg.occupancy_tables[k]):
                assert torch.equal(other, cg.occupancy_tables[k]), f"tags of table {k} differ across ranks"
        # (1b) the loser store (sharded over the ranks, read over NVLink): a forward over un-cached ids of this window
        # returns exactly their master rows, whichever rank holds them
        if rec.L is not None and tr.own_losers:
            # the rank-private loser list is exactly: ids of this rank's own batches of the window that are not cached
            loc_w = g.view(T, L, world, lb)[:, :, rank].reshape(T, L * lb)
            for k in range(T):
                u = torch.unique(loc_w[k])
                tags = cg.occupancy_tables[k]
                cached = (tags[u % tags.shape[0]] == u[:, None]).any(1)
                assert torch.equal(rec.loser_list(k), u[~cached]), f"window {w}: own loser list of table {k} is wrong"
        if rec.L is not None:
            probe = torch.zeros(T, lb, dtype=torch.int64, device=dev)
            for k in range(T):
                lo = rec.loser_list(k)
                if rec.L[k]:
                    probe[k] = lo[torch.arange(lb, device=dev) * max(rec.L[k] // lb, 1) % rec.L[k]]
            with torch.no_grad():
                outs, _ = cg(lS_o, probe, master, dev.index)
                cg.join_forward()
            torch.cuda.synchronize()
            for k in range(T):
                if rec.L[k]:
                    want = master.emb_l[k].weight.data[probe[k].cpu()]
                    assert torch.equal(outs[k].cpu(), want), f"window {w}: loser rows of table {k} differ from the master"
            checked_losers += sum(1 for k in range(T) if rec.L[k])
        dist.barrier()
        nxt = sg.window_ids(w + 1, L)
        submit(w + 1, nxt)
        loc = g.view(T, L, world, lb)[:, :, rank].reshape(T, L * lb).contiguous()
        X, Y = sl.dense_and_labels(w, L)
        for b in range(L):
            E, _ = tr.step(X[b * lb:(b + 1) * lb], lS_o, loc[:, b * lb:(b + 1) * lb], Y[b * lb:(b + 1) * lb])
            losses.append(float(E.item()))
            if j > 0 and j % agg == 0:
                torch.cuda.synchronize()
                before = [gathered(e.weight.data) for e in cg.emb_l]
                dirty = cg.dirty_bitmap().clone()
                alld = gathered(dirty)
                union = alld[0].clone()
                for o in alld[1:]:
                    union |= o
                broadcast_and_aggregate(cg, None, rank, args.table_agg_op)
                torch.cuda.synchronize()
                woff = 0
                for k in range(T):
                    rows = cg._cache_rows[k]
                    words = (rows + 31) // 32
                    bits = union[woff:woff + words].cpu().numpy().astype(np.uint32)
                    woff += words
                    touched = np.unpackbits(bits.view(np.uint8), bitorder="little")[:rows].astype(bool)
                    touched_t = torch.from_numpy(touched).to(dev)
                    mean = torch.stack(before[k]).sum(0) / world                      # (3) mean of the ranks
                    got = cg.emb_l[k].weight.data
                    torch.testing.assert_close(got[touched_t], mean[touched_t], rtol=1e-6, atol=1e-7)
                    assert torch.equal(got[~touched_t], before[k][rank][~touched_t])   # untouched rows untouched
                    nonaux = cg.cache_sizes[k] * ways
                    for other in gathered(got[:nonaux]):                              # (2) identical replicas
                        assert torch.equal(other, got[:nonaux]), f"table {k}: replicas differ after aggregation"
            j += 1
        g = nxt
    tr._plan_thread.join()
    # (4) the last boundary's write-back reached the master
    rec = tr._installed
    rec.wb_done.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    eo = np.concatenate([[0], np.cumsum(rec.E)])
    for k in range(T):
        if not rec.E[k]:
            continue
        ids, _slots, prim = rec.evict_list(k)
        ids, prim = ids.cpu(), prim.cpu().bool()
        rows = rec.evict_stage[eo[k]:eo[k] + rec.E[k]].cpu()
        # (every rank checks every row: with the write-back shared by the ranks, each row was written by ONE of them
        # into the host master that all of them map)
        assert torch.equal(master.emb_l[k].weight.data[ids[prim]], rows[prim]), f"write-back of table {k} missing"
    # (5) dense weights in sync, loss sane
    for seq in (tr.dlrm.bot_l, tr.dlrm.top_l):
        for layer in seq:
            if isinstance(layer, torch.nn.Linear):
                for other in gathered(layer.weight.data):
                    torch.testing.assert_close(other, layer.weight.data, rtol=1e-6, atol=1e-7)
    assert all(np.isfinite(losses)), losses
    cg.check_device_flags()
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check OK: world={world}, {n_windows} windows x {L} steps, loss {losses[0]:.4f} -> {losses[-1]:.4f}; "
              f"loser store {'sharded over the ranks' if tr.sharded_losers else ('own ids per rank' if tr.own_losers else 'all ids per rank')}, {checked_losers} table probes, "
              f"window scan {'sharded (marker)' if use_marker else 'whole window per rank'}, "
              f"write-back {'shared by the ranks' if tr.wb_sharded else 'by rank 0'}, "
              f"fill prefetch {'sharded over the ranks' if tr.sharded_fills else 'whole list per rank'}, "
              f"loss digest {hash(tuple(losses)) & 0xffffffff:08x}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
