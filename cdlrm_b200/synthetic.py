"""Criteo-shaped synthetic batches (the reference's random mode is unusable with its own
main).  Batch format of data_loader_terabyte.py:68-87:
X[B,13] = log(1 + U{0..100}), lS_o[T,B] = arange(B), lS_i[T,B] int64, T[B,1] ~ Bernoulli(0.25).

Index stream per table k: a bounded power law over ranks 1..n_k (exponent ``zipf_a``) or
uniform, scrambled over the id space with a multiplicative hash.
"""
import numpy as np
import torch

# cardinalities are external knowledge (not in the reference): Criteo Terabyte with the
# MLPerf 40M row cap, and Criteo Kaggle.
TERABYTE_ROWS = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208,
                 11938, 155, 4, 976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]
KAGGLE_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
               5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


class SyntheticStream:
    """Device-side twin of the loaders below for measurement at full size: ids come from cdlrm_synth_ids (one
    counter-based stream per table, csrc/synth.cu), so the id of (table, global step, sample) is a pure function
    of the seed.  ``ids`` generates any (step range x sample range) slice -- a rank's own training batches, or
    chunks of the GLOBAL window for the look-ahead planner -- and no rank ever has to hold the whole
    [T, lookahead x global batch] window."""

    def __init__(self, ln_emb, batch, device, dist="zipf", zipf_a=1.05, seed=123, dense_dim=13):
        self.ln = [int(n) for n in ln_emb]
        self.B = int(batch)                  # samples per step of the stream (the GLOBAL batch)
        self.dev = torch.device(device)
        self.dist, self.a, self.seed, self.dense_dim = dist, float(zipf_a), int(seed), dense_dim

    def _gen(self, w):
        g = torch.Generator(device=self.dev)
        g.manual_seed(self.seed * 1000003 + w)
        return g

    def ids(self, step0, n_steps, b0=0, nb=None, out=None, stream=None):
        """int64 [T, n_steps*nb] on the device: samples [b0, b0+nb) of steps [step0, step0+n_steps)."""
        import ctypes

        from ._lib import check, i64_array, lib
        nb = self.B - b0 if nb is None else int(nb)
        T, n = len(self.ln), int(n_steps) * nb
        if out is None:
            out = torch.empty(T, n, dtype=torch.int64, device=self.dev)
        assert out.is_cuda and out.dtype == torch.int64 and out.shape[0] == T and out.stride(1) == 1 and out.shape[1] >= n
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        check(lib.cdlrm_synth_ids(self.dev.index, 0, T, i64_array(self.ln), self.seed, self.B, int(step0), int(n_steps),
                                  int(b0), nb, int(self.dist == "uniform"), self.a, ctypes.c_void_p(out.data_ptr()),
                                  out.stride(0), ctypes.c_void_p(s.cuda_stream)))
        return out[:, :n]

    def window_ids(self, w, n_steps):
        """int64 [T, n_steps*B] on the device: the whole batch of steps [w*n_steps, (w+1)*n_steps)."""
        return self.ids(w * n_steps, n_steps)

    def dense_and_labels(self, w, n_steps, out=None):
        """X [N, dense_dim] = log(1 + U{0..100}), T [N, 1] ~ Bernoulli(0.25) for window w.  ``out=(X, T)``: generate
        in place into float32 buffers of those shapes (no allocation, no int64 temporary: a side stream that
        allocates gigabytes beside training stalls the host in cudaMalloc)."""
        g = self._gen(10_000_000 + w)
        N = n_steps * self.B
        if out is not None:
            X, T = out
            X.uniform_(0.0, 101.0, generator=g).floor_().clamp_(max=100.0).log1p_()
            T.uniform_(0.0, 1.0, generator=g)
            T.copy_(T < 0.25)                      # (a 25 MB bool temporary: small next to X)
            return X, T
        X = torch.log1p(torch.randint(0, 101, (N, self.dense_dim), generator=g, device=self.dev).float())
        T = (torch.rand(N, 1, generator=g, device=self.dev) < 0.25).float()
        return X, T


# ------------------------------------------------------------------------------------------
# host-side loaders: the role of dlrm_data_pytorch.make_criteo_data_and_loaders (:386-547),
# which hands main_no_ddp.py a (train_ld, test_ld, cache_ld) triple.  cache_ld is the
# Prefetcher's look-ahead twin of train_ld: it yields the SAME batches in the SAME order
# (cache_manager.py:87 reads only the sparse ids), which is the invariant the window FIFO
# relies on.
# ------------------------------------------------------------------------------------------


class SyntheticCriteoLoader:
    """Iterable of ``(X, lS_o, lS_i, T)`` CPU batches in the collate format of
    data_loader_terabyte.py:68-87.  Batch j is a pure function of (seed, split, j): any number
    of loaders built with the same arguments replay the same stream, whole epochs repeat it."""

    def __init__(self, ln_emb, batch, num_batches, dense_dim=13, dist="zipf", zipf_a=1.05, seed=123, split=0):
        self.ln = [int(n) for n in ln_emb]
        self.B, self.n = int(batch), int(num_batches)
        self.dense_dim, self.dist, self.a, self.seed, self.split = int(dense_dim), dist, float(zipf_a), int(seed), split
        self.lS_o = torch.arange(self.B).reshape(1, -1).repeat(len(self.ln), 1)

    def __len__(self):
        return self.n

    def sparse_ids(self, j, rng=None):
        rng = rng if rng is not None else np.random.default_rng([self.seed, self.split, j, 1])
        out = np.empty((len(self.ln), self.B), dtype=np.int64)
        for k, n in enumerate(self.ln):
            u = rng.random(self.B)
            if self.dist == "uniform" or n == 1:
                r = np.minimum((u * n).astype(np.int64), n - 1)
            else:  # inverse CDF of the continuous power law on [1, n+1), as SyntheticStream
                e = 1.0 - self.a
                r = np.clip((((n + 1.0) ** e - 1.0) * u + 1.0) ** (1.0 / e) - 1, 0, n - 1).astype(np.int64)
            out[k] = (r * 2654435761 + 40503 * k) % n
        return out

    def batch(self, j):
        ids = self.sparse_ids(j)
        rng = np.random.default_rng([self.seed, self.split, j, 2])
        X = np.log1p(rng.integers(0, 101, size=(self.B, self.dense_dim))).astype(np.float32)
        T = (rng.random((self.B, 1)) < 0.25).astype(np.float32)
        return torch.from_numpy(X), self.lS_o, torch.from_numpy(ids), torch.from_numpy(T)

    def __iter__(self):
        for j in range(self.n):
            yield self.batch(j)


def make_synthetic_data_and_loaders(args, ln_emb, m_den):
    """(train_ld, test_ld, cache_ld) for ``--data-generation synthetic`` (and ``random``: the reference's own
    random branch, main_no_ddp.py:539-547, builds no cache_ld / test_ld and cannot run).  Number of train
    batches: ``--num-batches`` or ``--data-size // --mini-batch-size`` (dlrm_data_pytorch.py:575)."""
    nb = int(args.num_batches) if args.num_batches > 0 else max(1, int(args.data_size) // int(args.mini_batch_size))
    kw = dict(dense_dim=int(m_den), dist=getattr(args, "synthetic_dist", "zipf"),
              zipf_a=getattr(args, "synthetic_zipf_a", 1.05), seed=int(args.numpy_rand_seed))
    train_ld = SyntheticCriteoLoader(ln_emb, args.mini_batch_size, nb, split=0, **kw)
    cache_ld = SyntheticCriteoLoader(ln_emb, args.mini_batch_size, nb, split=0, **kw)
    tb = args.test_mini_batch_size if getattr(args, "test_mini_batch_size", -1) > 0 else args.mini_batch_size
    test_ld = SyntheticCriteoLoader(ln_emb, tb, max(1, min(4, nb)), split=1, **kw)
    return train_ld, test_ld, cache_ld
