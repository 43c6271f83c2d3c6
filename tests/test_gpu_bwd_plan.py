"""Backward plan (the per-table stable sort of (slot, position) pairs that de-duplicates the sparse update,
autograd of model_no_ddp.py:200-202 + main_no_ddp.py:413): the thread-block-cluster radix sort against numpy's
stable argsort and, bit for bit, against the one-CTA-per-table kernel, over batch sizes that exercise partial
chunks, every cluster size, several sub-batches and unresolved positions (slot -1)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _plan(cg, slots, cluster):
    from cdlrm_b200._lib import check, lib
    vp = ctypes.c_void_p
    T, n = slots.shape
    check(lib.cdlrm_embed_set_option(0, cluster))
    try:
        nbytes = lib.cdlrm_embed_bwd_plan_bytes(T, n)
        buf = torch.zeros(nbytes + 256, dtype=torch.uint8, device=DEV)
        base = (buf.data_ptr() + 255) // 256 * 256
        check(lib.cdlrm_embed_bwd_plan(cg._ctx, 0, T, vp(slots.data_ptr()), slots.stride(0), n, vp(base),
                                       vp(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    finally:
        check(lib.cdlrm_embed_set_option(0, -1))
    off = base - buf.data_ptr()
    a = (T * n * 4 + 255) // 256 * 256
    pos = buf[off:off + T * n * 4].view(torch.int32).view(T, n).cpu().numpy()
    key = buf[off + a:off + a + T * n * 4].view(torch.int32).view(T, n).cpu().numpy().view(np.uint32)
    return key, pos


@pytest.mark.parametrize("n", [1, 33, 512, 700, 2049, 8192, 16384, 20000])
def test_cluster_sort_is_the_stable_sort(n):
    from cdlrm_b200 import model_no_ddp as M
    ln_emb = np.asarray([3, 155, 5000, 2_500_000])
    cg = M.Embedding_Table_Cache_Group(16, ln_emb, 150000, 64, 16).to(DEV)
    cg._ensure_ctx(None)
    rng = np.random.default_rng(n)
    T = len(ln_emb)
    slots = np.empty((T, n), dtype=np.int32)
    for k in range(T):
        rows = cg._cache_rows[k]
        z = (rng.zipf(1.2, size=n) * 7919) % rows                     # heavy duplicates
        u = rng.integers(0, rows, size=n)
        slots[k] = np.where(rng.random(n) < 0.5, z, u)
        slots[k, rng.random(n) < 0.02] = -1                           # unresolved positions
    d_slots = torch.from_numpy(slots).to(DEV)
    ref_key, ref_pos = _plan(cg, d_slots, 0)                          # one CTA per table
    SM = 16384                                                        # CDLRM_SORT_MAX: sub-batches are sorted separately
    for k in range(T):
        rows = cg._cache_rows[k]
        for j0 in range(0, n, SM):
            sl = slots[k, j0:j0 + SM].astype(np.int64)
            keyed = np.where((sl < 0) | (sl >= rows), rows, sl)
            order = np.argsort(keyed, kind="stable")
            assert np.array_equal(ref_pos[k, j0:j0 + SM], j0 + order)
            assert np.array_equal(ref_key[k, j0:j0 + SM], keyed[order].astype(np.uint32))
    for cluster in (1, 2, 4, 8):
        key, pos = _plan(cg, d_slots, cluster)
        assert np.array_equal(key, ref_key) and np.array_equal(pos, ref_pos), f"cluster size {cluster}"
