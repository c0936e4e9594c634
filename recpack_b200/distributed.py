"""Multi-GPU plumbing of the hot path: one process per GPU (torchrun), item rows sharded for fit,
users sharded for scoring.  The only data-path exchange is one all-gather of the pruned similarity
lists (I x K x 12 B in total) and one all-reduce of the metric sums; both go through
torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(weights, parts: int):
    """Cut points of `parts` contiguous shards with balanced total weight: cuts[r]..cuts[r+1]."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(c, c[-1] * p / parts)) for p in range(parts + 1)]
    cuts[0], cuts[-1] = 0, len(w)
    for p in range(1, parts + 1):  # monotone even with zero weights
        cuts[p] = max(cuts[p], cuts[p - 1])
    return cuts


def fit_work_per_item(X):
    """Work of item row i in the sparse Gram: sum of the history lengths of its users (+ a constant
    for the row's fixed cost)."""
    d = np.diff(X.indptr).astype(np.float64)
    return np.bincount(X.indices, weights=np.repeat(d, np.diff(X.indptr)), minlength=X.shape[1]) + 2000.0


def score_work_per_user(X, K: int):
    return np.diff(X.indptr).astype(np.float64) * K + 3.0 * X.shape[1] / 64


class ShardExchange:
    """Pre-allocated buffers for the all-gather of per-rank top-K lists into full [I, K] arrays."""

    def __init__(self, cuts, K, device, dist):
        import torch

        self.cuts, self.K, self.dist = list(cuts), int(K), dist
        self.world = len(cuts) - 1
        self.rank = dist.get_rank()
        self.maxrows = max(cuts[r + 1] - cuts[r] for r in range(self.world))
        I = cuts[-1]
        mk = lambda shape, dt, fill: torch.full(shape, fill, dtype=dt, device=device)
        self.p_idx = mk((self.maxrows, K), torch.int32, -1)
        self.p_val = mk((self.maxrows, K), torch.float64, 0)
        self.p_len = mk((self.maxrows,), torch.int32, 0)
        self.g_idx = mk((self.world, self.maxrows, K), torch.int32, -1)
        self.g_val = mk((self.world, self.maxrows, K), torch.float64, 0)
        self.g_len = mk((self.world, self.maxrows), torch.int32, 0)
        self.p_ent = mk((self.maxrows, K), torch.int64, -1)  # packed model rows (uint64 bit patterns), all-ones = unused
        self.g_ent = mk((self.world, self.maxrows, K), torch.int64, -1)
        self.vmax = mk((1,), torch.float64, 0)  # largest similarity of the model (device scalar)
        self.all_idx = mk((I, K), torch.int32, -1)
        self.all_val = mk((I, K), torch.float64, 0)
        self.all_len = mk((I,), torch.int32, 0)

    def local_out(self):
        """Views of this rank's send buffers shaped like its fit output: pass them as `out=` of the fit."""
        rows = self.cuts[self.rank + 1] - self.cuts[self.rank]
        return {"idx": self.p_idx[:rows], "cnt": None, "val": self.p_val[:rows], "len": self.p_len[:rows]}

    def row_source(self):
        """int64[I]: row of the gathered [world * maxrows, K] arrays that holds item row i."""
        if getattr(self, "_row_src", None) is None:
            import torch

            src = np.empty(self.cuts[-1], dtype=np.int64)
            for r in range(self.world):
                b, e = self.cuts[r], self.cuts[r + 1]
                src[b:e] = r * self.maxrows + np.arange(e - b)
            self._row_src = torch.from_numpy(src).to(self.g_idx.device)
        return self._row_src

    def gather_padded(self):
        """All-gather of the send buffers (filled through local_out()).  Returns the padded arrays
        ([world * maxrows, K], [world * maxrows]) -- rpk_model_load_topk_rows reads them through row_source()."""
        self.dist.all_gather_into_tensor(self.g_idx.view(-1, self.K), self.p_idx)
        self.dist.all_gather_into_tensor(self.g_val.view(-1, self.K), self.p_val)
        self.dist.all_gather_into_tensor(self.g_len.view(-1), self.p_len)
        return self.g_idx.view(-1, self.K), self.g_val.view(-1, self.K), self.g_len.view(-1)

    def gather_packed(self, engine=None):
        """Exchange in the model's own format: this rank's lists (filled through local_out()) are packed into
        model rows (rpk_model_pack_rows; 8 bytes per entry, already in column order) and all-gathered together with
        the row lengths.  Returns ([world * maxrows, K] int64, [world * maxrows] int32) for
        rpk_model_load_packed_rows_v + row_source(); the model's largest value is left in self.vmax (device).  With
        engine=None the caller has filled p_ent (and vmax) itself."""
        rows = self.cuts[self.rank + 1] - self.cuts[self.rank]
        if engine is not None:
            # one fixed-point scale for the whole model: the largest similarity of any rank, all-reduced on the stream
            # (a device scalar: no host round trip anywhere in the exchange)
            engine.model_vmax(self.K, self.p_val[:rows], self.p_len[:rows], self.vmax)
            self.dist.all_reduce(self.vmax, op=self.dist.ReduceOp.MAX)
            engine.model_pack_rows_v(self.cuts[-1], self.K, self.p_idx[:rows], self.p_val[:rows], self.p_len[:rows], self.vmax,
                                     self.p_ent[:rows])
        self.dist.all_gather_into_tensor(self.g_ent.view(-1, self.K), self.p_ent)
        self.dist.all_gather_into_tensor(self.g_len.view(-1), self.p_len)
        return self.g_ent.view(-1, self.K), self.g_len.view(-1)

    def gather(self, idx, val, ln):
        """idx/val/ln: this rank's rows (cuts[rank]..cuts[rank+1]).  Returns the full, unpadded arrays."""
        rows = self.cuts[self.rank + 1] - self.cuts[self.rank]
        self.p_idx[:rows].copy_(idx)
        self.p_val[:rows].copy_(val)
        self.p_len[:rows].copy_(ln)
        g_idx, g_val, g_len = self.gather_padded()
        src = self.row_source()
        self.all_idx.copy_(g_idx[src])
        self.all_val.copy_(g_val[src])
        self.all_len.copy_(g_len[src])
        return self.all_idx, self.all_val, self.all_len
