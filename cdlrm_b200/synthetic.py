"""Criteo-shaped synthetic batches (the reference's random mode is unusable with its own
main).  Batch format of data_loader_terabyte.py:68-87:
X[B,13] = log(1 + U{0..100}), lS_o[T,B] = arange(B), lS_i[T,B] int64, T[B,1] ~ Bernoulli(0.25).

Index stream per table k: a bounded power law over ranks 1..n_k (exponent ``zipf_a``) or
uniform, scrambled over the id space with a multiplicative hash; generated on the device
in whole windows so that the look-ahead planner and the training steps read the same ids.
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
    def __init__(self, ln_emb, batch, device, dist="zipf", zipf_a=1.05, seed=123, dense_dim=13):
        self.ln = [int(n) for n in ln_emb]
        self.B = int(batch)
        self.dev = torch.device(device)
        self.dist, self.a, self.seed, self.dense_dim = dist, float(zipf_a), int(seed), dense_dim

    def _gen(self, w):
        g = torch.Generator(device=self.dev)
        g.manual_seed(self.seed * 1000003 + w)
        return g

    def window_ids(self, w, n_steps):
        """int64 [T, n_steps*B] on the device: ids of steps [0, n_steps) of window w."""
        N = n_steps * self.B
        g = self._gen(w)
        out = torch.empty(len(self.ln), N, dtype=torch.int64, device=self.dev)
        for k, n in enumerate(self.ln):
            u = torch.rand(N, generator=g, device=self.dev, dtype=torch.float64)
            if self.dist == "uniform" or n == 1:
                r = (u * n).long().clamp_(max=n - 1)
            else:  # inverse CDF of the continuous power law on [1, n+1)
                e = 1.0 - self.a
                r = (((n + 1.0) ** e - 1.0) * u + 1.0).pow_(1.0 / e).long().sub_(1).clamp_(0, n - 1)
            out[k] = (r * 2654435761 + 40503 * k) % n
        return out

    def dense_and_labels(self, w, n_steps):
        g = self._gen(10_000_000 + w)
        N = n_steps * self.B
        X = torch.log1p(torch.randint(0, 101, (N, self.dense_dim), generator=g, device=self.dev).float())
        T = (torch.rand(N, 1, generator=g, device=self.dev) < 0.25).float()
        return X, T
