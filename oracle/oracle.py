"""CPU oracle for the cDLRM look-ahead embedding-cache hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-numpy restatement of the
reference's algorithm (lkp411/cDLRM, files cited per function as
``file:line`` relative to the reference tree).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline -- never as part of the product path.  The product
(``cdlrm_b200``) fails loudly when its CUDA library is missing; it never
falls back to this code.

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference ITSELF, run in the build container by ``oracle/gen_golden.py``
(which imports ``/root/reference`` unmodified through a CPU-device shim) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
every function below against those vectors.

The arithmetic the reference delegates to PyTorch (``torch.unique``,
``torch.remainder``, ``torch.distributions.Categorical``, ``nn.EmbeddingBag``,
``torch.optim.SGD``, ``torch.bmm``) is restated here from its observed
behaviour under torch 2.11.0 CPU (the image's version; the reference pins
none):

* ``Categorical(avail).sample()`` == ``argmax(probs / q)`` with
  ``q = empty(rows, ways).exponential_(1)`` drawn from the global CPU
  mt19937 generator, first index winning ties (``torch.multinomial`` fast
  path).
* ``exponential_(1)`` on a float32 CPU tensor == ``float32(-log1p(-u))`` with
  ``u = (r64 & (2**53-1)) * 2**-53`` and ``r64 = (mt32() << 32) | mt32()``,
  mt19937 seeded with ``init_genrand(seed)``; element order row-major; the
  stream is split-invariant and thread-count-invariant.
* duplicate ``index_put_`` on CPU with one thread: last write wins.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# geometry (model_no_ddp.py:122-125, 319-331)
# --------------------------------------------------------------------------


def is_prime_ref(n: int) -> bool:
    """model_no_ddp.py:319-331 -- NOT a real primality test.

    Only odd/even divisors ``3 <= i`` with ``i*i < n`` are tried (so even
    numbers such as 10006 = 2*5003 and perfect squares of primes pass)."""
    if n == 1 or n == 2:
        return False
    i = 3
    while i * i < n:
        if n % i == 0:
            return False
        i += 1
    return True


def find_next_prime(max_cache_size: int):
    """model_no_ddp.py:122-125."""
    for i in range(max_cache_size, 2 * max_cache_size):
        if is_prime_ref(i):
            return i
    return None


# --------------------------------------------------------------------------
# torch CPU generator restatement (used by main_no_ddp.py:183-185 through
# torch.distributions.Categorical -> torch.multinomial -> exponential_)
# --------------------------------------------------------------------------


class TorchCpuGenerator:
    """mt19937 stream identical to ``torch.manual_seed(seed)``'s CPU generator
    for the ``exponential_`` draws consumed by the victim-way sampler."""

    def __init__(self, seed: int):
        # RandomState(int) uses init_genrand(seed) == torch's mt19937(seed).
        self._rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)
        self.draws = 0

    def random64(self, n: int) -> np.ndarray:
        raw = self._rs.randint(0, 2 ** 32, size=2 * n, dtype=np.uint64)
        return (raw[0::2] << np.uint64(32)) | raw[1::2]

    def exponential_f32(self, n: int) -> np.ndarray:
        """``torch.empty(n).exponential_(1)`` (float32)."""
        if n == 0:
            return np.empty(0, dtype=np.float32)
        r = self.random64(n)
        u = (r & np.uint64((1 << 53) - 1)).astype(np.float64) * (1.0 / (1 << 53))
        self.draws += n
        return (-np.log1p(-u)).astype(np.float32)


def sample_ways(avail: np.ndarray, q: np.ndarray) -> np.ndarray:
    """main_no_ddp.py:183-185.  ``avail`` [rows, ways] bool, ``q`` [rows, ways]
    float32 exponential draws.  Returns int64 way per row."""
    if avail.shape[0] == 0:
        return np.empty(0, dtype=np.int64)
    p = avail.astype(np.float32)
    p = p / p.sum(axis=1, keepdims=True, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = p / q
    return r.argmax(axis=1).astype(np.int64)


# --------------------------------------------------------------------------
# storage (model_no_ddp.py:101-147)
# --------------------------------------------------------------------------


class OracleCache:
    """Embedding_Table_Cache_Group state: tags (``occupancy_tables``) and cache
    rows (``emb_l[k].weight``), model_no_ddp.py:102-147."""

    def __init__(self, m_spa, ln_emb, max_cache_size, aux_table_size, num_ways,
                 init_weights=None):
        self.dim = int(m_spa)
        self.ln_emb = [int(n) for n in ln_emb]
        self.num_ways = int(num_ways)
        self.aux = int(aux_table_size)
        self.max_cache_size = find_next_prime(int(max_cache_size))
        self.cache_sizes = [n if n < self.max_cache_size else self.max_cache_size
                            for n in self.ln_emb]
        self.tags = [np.full((s, self.num_ways), -1, dtype=np.int64) for s in self.cache_sizes]
        self.weight = []
        for k, s in enumerate(self.cache_sizes):
            rows = self.num_ways * s + self.aux
            if init_weights is not None:
                w = np.array(init_weights[k], dtype=np.float32, copy=True)
                assert w.shape == (rows, self.dim)
            else:
                w = np.zeros((rows, self.dim), dtype=np.float32)
            self.weight.append(w)

    def set_indices(self, k, ids):
        """compute_set_indices, model_no_ddp.py:127-128 (torch.remainder)."""
        return np.remainder(ids, self.cache_sizes[k])


# --------------------------------------------------------------------------
# a1: Prefetcher.process_batch_slice (cache_manager.py:27-46)
# --------------------------------------------------------------------------


def process_batch_slice(window_ids, master):
    """``window_ids`` [T, N] int64; ``master`` list of [n_k, d] float32.
    Returns (rows_per_table, unique_per_table, maps_per_table)."""
    rows, uniq, maps = [], [], []
    for k in range(len(master)):
        u = np.unique(window_ids[k])                       # :32 sorted ascending
        m = np.full((int(u.max()) + 1, 1), -1, dtype=np.int64)   # :36-39
        m[u, 0] = np.arange(u.shape[0])
        uniq.append(u)
        maps.append(m)
        rows.append(master[k][u])                          # :44 -> model_no_ddp.py:80-87
    return rows, uniq, maps


# --------------------------------------------------------------------------
# a5: CacheEmbeddings (main_no_ddp.py:148-209)
# --------------------------------------------------------------------------


def plan_table(tags, S, ways, u, draw_q):
    """Integer part of CacheEmbeddings for one table.  Mutates ``tags``.

    Returns dict with the per-table decisions (all int64):
      n_hit, n_dropped, surv_ids [R], surv_sets [R], way [R],
      evict_ids [E], evict_slots [E], slot [R] (= S*way+set)."""
    s = np.remainder(u, S)                                            # :155
    match = tags[s] == u[:, None]                                     # :160
    hit = match.any(axis=1)
    hit_pos = np.nonzero(hit)[0]
    miss_pos = np.nonzero(~hit)[0]
    hit_sets = s[hit_pos]                                             # :164
    hit_ways = np.nonzero(match)[1]                                   # :165
    need_u = u[miss_pos]                                              # :167
    need_s = s[miss_pos]
    avail = np.ones(tags.shape, dtype=bool)                           # :171
    avail[hit_sets, hit_ways] = False                                 # :172
    full = ~avail.any(axis=1)                                         # :173 (as a mask over sets)
    keep = np.nonzero(~full[need_s])[0]                               # :176-177
    n_dropped = need_u.shape[0] - keep.shape[0]
    need_u = need_u[keep]                                             # :179-180
    need_s = need_s[keep]
    R = need_u.shape[0]
    q = draw_q(R * ways).reshape(R, ways)
    way = sample_ways(avail[need_s], q)                               # :183-185
    old = tags[need_s, way]                                           # :190
    ev = np.nonzero(old != -1)[0]
    evict_ids = old[ev]                                               # :196
    evict_slots = S * way[ev] + need_s[ev]                            # :194
    slot = S * way + need_s                                           # :203
    # :204 -- sequential (single-thread) index_put_: last write wins (numpy's
    # documented rule for repeated indices in an assignment is the same).
    tags[need_s, way] = need_u
    return dict(n_hit=int(hit_pos.shape[0]), n_dropped=int(n_dropped), surv_ids=need_u,
                surv_sets=need_s, way=way, evict_ids=evict_ids, evict_slots=evict_slots,
                slot=slot, q=q)


def cache_embeddings(rows, uniq, maps, cache: OracleCache, draw_q):
    """main_no_ddp.py:148-209.  Returns (eviction_data, plans); eviction_data is
    the list the reference puts on ``eviction_fifo`` (:208-209)."""
    eviction_data, plans = [], []
    for k, table_cache in enumerate(rows):
        S = cache.cache_sizes[k]
        p = plan_table(cache.tags[k], S, cache.num_ways, uniq[k], draw_q)
        evict_rows = cache.weight[k][p["evict_slots"]].copy()          # :197
        eviction_data.append((p["evict_ids"], evict_rows))             # :199
        src = maps[k][p["surv_ids"]].ravel()                           # :205
        # :206 -- last write wins (the contract; the reference's CUDA
        # index_put_ is unordered, its single-thread CPU run is last-wins).
        cache.weight[k][p["slot"]] = table_cache[src]
        plans.append(p)
    return eviction_data, plans


# --------------------------------------------------------------------------
# a3: eviction_manager body (cache_manager.py:57-62)
# --------------------------------------------------------------------------


def eviction_writeback(master, eviction_data, average_on_writeback=False):
    for k, (ids, emb) in enumerate(eviction_data):
        if ids.shape[0] == 0:
            continue
        if average_on_writeback:
            master[k][ids] = (master[k][ids] + emb) / 2
        else:
            master[k][ids] = emb


# --------------------------------------------------------------------------
# a8: Embedding_Table_Cache_Group.forward (model_no_ddp.py:149-212)
# --------------------------------------------------------------------------


def forward_table(cache: OracleCache, k, offsets, ids, master_k):
    S, ways = cache.cache_sizes[k], cache.num_ways
    tags, w = cache.tags[k], cache.weight[k]
    s = np.remainder(ids, S)                                            # :166
    match = tags[s] == ids[:, None]                                     # :168
    hit = match.any(axis=1)
    hit_pos = np.nonzero(hit)[0]
    miss_pos = np.nonzero(~hit)[0]
    hit_ways = np.nonzero(match[hit_pos])[1]                            # :173
    slots = np.empty(ids.shape[0], dtype=np.int64)
    slots[hit_pos] = S * hit_ways + s[hit_pos]                          # :174
    missing = ids[miss_pos]
    aux = S * ways + np.arange(missing.shape[0], dtype=np.int64)        # :177
    if aux.shape[0] > cache.aux:
        raise IndexError("aux (victim) region overflow")               # index error in the reference
    w[aux] = master_k[missing]                                          # :179
    slots[miss_pos] = aux                                               # :185
    # EmbeddingBag(mode="sum") :202
    nb = offsets.shape[0]
    out = np.zeros((nb, cache.dim), dtype=np.float32)
    ends = np.append(offsets[1:], ids.shape[0])
    for b in range(nb):
        lo, hi = int(offsets[b]), int(ends[b])
        if hi - lo == 1:
            out[b] = w[slots[lo]]
        else:
            for j in range(lo, hi):
                out[b] += w[slots[j]]
    return out, slots.astype(np.int32), int(missing.shape[0])          # :204 (.int())


def forward(cache: OracleCache, lS_o, lS_i, master):
    ly, idxs, n_miss = [], [], []
    for k in range(len(cache.weight)):
        o, s, m = forward_table(cache, k, np.asarray(lS_o[k]), np.asarray(lS_i[k]), master[k])
        ly.append(o)
        idxs.append(s)
        n_miss.append(m)
    return ly, idxs, n_miss


def forward_table_fast(cache: OracleCache, k, ids, master_k):
    """Vectorised P=1 variant of forward_table (same results); used by the
    timed cpu_baseline leg so the port is not handicapped by Python loops."""
    S, ways = cache.cache_sizes[k], cache.num_ways
    tags, w = cache.tags[k], cache.weight[k]
    s = np.remainder(ids, S)
    match = tags[s] == ids[:, None]
    hit = match.any(axis=1)
    slots = S * match.argmax(axis=1) + s
    miss_pos = np.nonzero(~hit)[0]
    aux = S * ways + np.arange(miss_pos.shape[0], dtype=np.int64)
    w[aux] = master_k[ids[miss_pos]]
    slots[miss_pos] = aux
    return w[slots], slots.astype(np.int32), int(miss_pos.shape[0])


# --------------------------------------------------------------------------
# a9: EmbeddingBag backward + SGD (autograd; main_no_ddp.py:376,409,413)
# --------------------------------------------------------------------------


def backward_sgd_table(w, slots, offsets, dV, lr):
    """``weight[slot] += -lr * dV[bag(slot)]`` accumulated in index order
    (torch CPU sparse add is sequential over nnz)."""
    n = slots.shape[0]
    ends = np.append(offsets[1:], n)
    bag = np.repeat(np.arange(offsets.shape[0]), (ends - offsets).astype(np.int64))
    g = (-np.float32(lr)) * dV[bag]
    np.add.at(w, slots.astype(np.int64), g)


# --------------------------------------------------------------------------
# a11: DLRM_Net.interact_features, "dot" (model_no_ddp.py:272-293)
# --------------------------------------------------------------------------


def tril_pairs(nf, itself=False):
    """:288-291 -- row-major strict (or inclusive) lower triangle."""
    off = 1 if itself else 0
    li = [i for i in range(nf) for _ in range(i + off)]
    lj = [j for i in range(nf) for j in range(i + off)]
    return np.asarray(li, dtype=np.int64), np.asarray(lj, dtype=np.int64)


def interact_fwd(x, ly, itself=False):
    B, d = x.shape
    T = np.concatenate([x] + list(ly), axis=1).reshape(B, -1, d)        # :276
    Z = np.einsum("bid,bjd->bij", T.astype(np.float64), T.astype(np.float64))  # :278 (fp64 accumulate)
    li, lj = tril_pairs(T.shape[1], itself)
    return np.concatenate([x, Z[:, li, lj].astype(np.float32)], axis=1)  # :293


def interact_bwd(x, ly, dR, itself=False):
    """Gradient of interact_fwd w.r.t. x and every ly[k] (autograd of :276-293)."""
    B, d = x.shape
    T = np.concatenate([x] + list(ly), axis=1).reshape(B, -1, d).astype(np.float64)
    nf = T.shape[1]
    li, lj = tril_pairs(nf, itself)
    dZ = np.zeros((B, nf, nf), dtype=np.float64)
    dZ[:, li, lj] = dR[:, d:]
    dT = np.einsum("bij,bjd->bid", dZ + dZ.transpose(0, 2, 1), T)
    dT[:, 0, :] += dR[:, :d]
    dT = dT.astype(np.float32)
    return dT[:, 0, :], [dT[:, k + 1, :] for k in range(nf - 1)]


# --------------------------------------------------------------------------
# a10: broadcast_and_aggregate (main_no_ddp.py:250-292)
# --------------------------------------------------------------------------


def aggregate(weights_per_rank, idxs_per_rank, reduce_op="mean"):
    """``weights_per_rank[r][k]`` cache rows of rank r; ``idxs_per_rank[r]``
    int32 [T, n] slot lists of rank r.  In-place on every rank's weights."""
    W = len(weights_per_rank)
    lookups = np.concatenate(idxs_per_rank, axis=1)                     # :268
    for k in range(lookups.shape[0]):
        u = np.unique(lookups[k]).astype(np.int64)                      # :270
        if reduce_op == "mean":
            sl = [weights_per_rank[r][k][u] / np.float32(W) for r in range(W)]   # :277
            red = sl[0].copy()
            for r in range(1, W):
                red = red + sl[r]
        elif reduce_op == "sum":
            red = weights_per_rank[0][k][u].copy()
            for r in range(1, W):
                red = red + weights_per_rank[r][k][u]
        elif reduce_op == "max":
            red = weights_per_rank[0][k][u].copy()
            for r in range(1, W):
                red = np.maximum(red, weights_per_rank[r][k][u])
        else:
            raise ValueError(reduce_op)
        for r in range(W):
            weights_per_rank[r][k][u] = red                             # :292


# --------------------------------------------------------------------------
# a12 + oracle schedule: one window on one rank
# --------------------------------------------------------------------------


def install_window(cache: OracleCache, master, window_ids, gen: TorchCpuGenerator,
                   average_on_writeback=False):
    """process_batch_slice -> CacheEmbeddings -> eviction write-back applied
    inline (sequential schedule).  Returns (eviction_data, plans, uniq)."""
    rows, uniq, maps = process_batch_slice(window_ids, master)
    ev, plans = cache_embeddings(rows, uniq, maps, cache, gen.exponential_f32)
    eviction_writeback(master, ev, average_on_writeback)
    return ev, plans, uniq


# --------------------------------------------------------------------------
# vectorised variants used ONLY by the timed cpu_baseline / --impl reference legs of
# bench.py (same results as the functions above; tests/test_oracle_golden.py checks that)
# --------------------------------------------------------------------------


def backward_sgd_table_fast(w, slots, dV, lr):
    """P=1 sparse SGD: duplicates merged with a stable sort + segmented sum, then one
    update per distinct slot (what torch's coalesce + add does)."""
    s = slots.astype(np.int64)
    order = np.argsort(s, kind="stable")
    ss = s[order]
    starts = np.flatnonzero(np.r_[True, ss[1:] != ss[:-1]])
    sums = np.add.reduceat(dV[order], starts, axis=0)
    w[ss[starts]] += (-np.float32(lr)) * sums


def interact_fwd_fast(x, ly, itself=False):
    """float32 batched sgemm like torch.bmm (model_no_ddp.py:276-293)."""
    B, d = x.shape
    T = np.concatenate([x] + list(ly), axis=1).reshape(B, -1, d)
    Z = np.matmul(T, T.transpose(0, 2, 1))
    li, lj = tril_pairs(T.shape[1], itself)
    return np.concatenate([x, Z[:, li, lj]], axis=1), T


def interact_bwd_fast(T, dR, itself=False):
    B, nf, d = T.shape
    li, lj = tril_pairs(nf, itself)
    dZ = np.zeros((B, nf, nf), dtype=np.float32)
    dZ[:, li, lj] = dR[:, d:]
    dT = np.matmul(dZ + dZ.transpose(0, 2, 1), T)
    dT[:, 0, :] += dR[:, :d]
    return dT
