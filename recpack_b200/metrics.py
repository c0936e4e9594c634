"""Top-K metrics on the GPU -- drop-ins for recpack.metrics.

With ``recpack`` importable every class here subclasses the reference's class of the same name
(recpack/metrics/{dcg,recall,precision,reciprocal_rank,hit,coverage}.py): ``name``, ``results``, ``value``,
``num_users`` / ``num_items`` and the constructors are inherited, ``calculate(y_true, y_pred)`` is replaced --
the ranking of the prediction rows (recpack/util.py:50-77) and the metric itself run in rpk_topk_csr /
rpk_metrics_topn / rpk_coverage_topn.  A prediction matrix produced by this package's
``predict(..., predict_topK=N)`` carries its rank-ordered lists and skips the ranking, as long as it has not
been edited since.

``y_true`` is used as a BINARY matrix (its sparsity structure; stored zeros are dropped), which is what the
pipeline passes (``test_data_out.binary_values``, pipelines/pipeline.py:160) and what the reference's formulas
assume (metrics/dcg.py:109-111)."""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

from . import _ref
from .engine import get_engine
from .matrix import device_structure, to_host
from .util import ranks_from_lists, top_k_lists

if _ref.HAVE_RECPACK:
    _m = _ref.ref_metrics
    _bases = {"NDCGK": _m.NDCGK, "DCGK": _m.DCGK, "RecallK": _m.RecallK, "CalibratedRecallK": _m.CalibratedRecallK,
              "PrecisionK": _m.PrecisionK, "ReciprocalRankK": _m.ReciprocalRankK, "HitK": _m.HitK, "CoverageK": _m.CoverageK}
else:
    from . import _mirror

    _bases = {k: _mirror.ListwiseMetricK for k in ("NDCGK", "DCGK", "RecallK", "CalibratedRecallK", "PrecisionK", "ReciprocalRankK")}
    _bases["HitK"] = _mirror.ElementwiseMetricK
    _bases["CoverageK"] = _mirror.GlobalMetricK


def _matrix_signature(M):
    from .base import matrix_signature

    return matrix_signature(M)


def _column_csr(vals: np.ndarray) -> csr_matrix:
    """``csr_matrix(vals.reshape(-1, 1))`` (the layout of the reference's ``scores_``: one row per evaluated user, zeros
    not stored) assembled directly instead of through scipy's dense -> COO -> CSR route."""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    nz = vals != 0
    indptr = np.zeros(vals.shape[0] + 1, dtype=np.int32)
    np.cumsum(nz, out=indptr[1:])
    data = vals[nz]
    out = csr_matrix((data, np.zeros(data.shape[0], dtype=np.int32), indptr), shape=(vals.shape[0], 1))
    out.has_canonical_format = True
    return out


class GpuTopKMixin:
    """``calculate`` of MetricTopK (metrics/base.py:172-193) on the GPU: drop users without true items, rank the
    K best stored predictions per user, evaluate."""

    _kind = None

    def _ranked_lists(self, y_true: csr_matrix, y_pred: csr_matrix, engine):
        """(idx, len) on the host plus (idx, len) where the metric kernels should read them."""
        K = int(self.K)
        lists = getattr(y_pred, "_rpk_topn", None)
        sig = getattr(y_pred, "_rpk_topn_sig", None)
        if (lists is not None and lists[0].shape[1] >= K and lists[0].shape[0] == y_true.shape[0]
                and (sig is None or sig == _matrix_signature(y_pred))):
            idx, ln = lists
            dev = getattr(y_pred, "_rpk_topn_dev", None)
            if dev is not None and dev[2] == engine.device:
                return idx, ln, dev[0], dev[1]  # the lists are still on the device: nothing to upload
        else:
            idx, ln = top_k_lists(y_pred, K)  # a foreign (or edited) matrix: rank its rows (rpk_topk_csr)
        return idx, ln, np.ascontiguousarray(idx), np.ascontiguousarray(ln)

    def _prepare(self, y_true, y_pred):
        y_true = csr_matrix(y_true) if not isinstance(y_true, csr_matrix) else y_true
        self._verify_shape(y_true, y_pred)
        engine = get_engine()
        idx, ln, top_idx, top_len = self._ranked_lists(y_true, y_pred, engine)
        yt, t_ptr, t_idx, t_ptr_d, t_idx_d = device_structure(y_true, engine.device)
        users = np.flatnonzero(np.diff(t_ptr) > 0)  # metrics/base.py:106-123
        self.user_id_map_ = users
        self.num_users_, self.num_items_ = len(users), yt.shape[1]
        self._lists = (idx, ln, users, y_true.shape)
        return engine, idx, top_idx, top_len, yt, t_ptr_d, t_idx_d, users

    @property
    def y_pred_top_K_(self):
        """Ranks 1..K of the recommended items per evaluated user (metrics/base.py:189), built on demand."""
        idx, ln, users, shape = self._lists
        K = int(self.K)
        return ranks_from_lists(idx[users, :K], np.minimum(ln[users], K), (len(users), shape[1]))

    @y_pred_top_K_.setter
    def y_pred_top_K_(self, value):  # the reference's calculate assigns it; ours derives it from the lists
        pass


class GpuListwiseMixin(GpuTopKMixin):
    def calculate(self, y_true: csr_matrix, y_pred: csr_matrix) -> None:
        engine, idx, top_idx, top_len, yt, t_ptr_d, t_idx_d, users = self._prepare(y_true, y_pred)
        sums, n_users, per_user = engine.metrics_topn(yt.shape[0], idx.shape[1], top_idx, top_len, t_ptr_d, t_idx_d,
                                                      [(self._kind, int(self.K))])
        if not isinstance(per_user, np.ndarray):
            engine.sync()
            (per_user,) = to_host(per_user)
        self.scores_ = _column_csr(per_user[0, users])
        self.sum_, self.n_users_ = float(sums[0]), int(n_users)  # device-side reduction (used by the sharded bench)

    @property
    def _indices(self):
        n = len(self.user_id_map_)
        return np.arange(n), np.zeros(n, dtype=np.int32)


class NDCGK(GpuListwiseMixin, _bases["NDCGK"]):
    """metrics/dcg.py:73-128."""

    _kind = "ndcg"


class DCGK(GpuListwiseMixin, _bases["DCGK"]):
    """metrics/dcg.py:21-52."""

    _kind = "dcg"


class RecallK(GpuListwiseMixin, _bases["RecallK"]):
    """metrics/recall.py:21-48."""

    _kind = "recall"


class CalibratedRecallK(GpuListwiseMixin, _bases["CalibratedRecallK"]):
    """metrics/recall.py:58-85."""

    _kind = "calibrated_recall"


class PrecisionK(GpuListwiseMixin, _bases["PrecisionK"]):
    """metrics/precision.py:12-50: hits / K (fewer than K recommendations count as misses)."""

    _kind = "precision"


class ReciprocalRankK(GpuListwiseMixin, _bases["ReciprocalRankK"]):
    """metrics/reciprocal_rank.py:13-40: 1 / rank of the first hit, 0 without one."""

    _kind = "reciprocal_rank"


class HitK(GpuTopKMixin, _bases["HitK"]):
    """metrics/hit.py:20-45: ``value`` = mean number of hits among the first K places; ``results`` lists every
    (user, recommended item) pair with 1 for a hit.  The per-user hit counts come from rpk_metrics_topn; the
    element-wise ``scores_`` matrix behind ``results`` is built on demand from the rank-ordered lists."""

    def calculate(self, y_true: csr_matrix, y_pred: csr_matrix) -> None:
        engine, idx, top_idx, top_len, yt, t_ptr_d, t_idx_d, users = self._prepare(y_true, y_pred)
        sums, n_users, per_user = engine.metrics_topn(yt.shape[0], idx.shape[1], top_idx, top_len, t_ptr_d, t_idx_d,
                                                      [("hits", int(self.K))])
        if not isinstance(per_user, np.ndarray):
            engine.sync()
            (per_user,) = to_host(per_user)
        self.hits_per_user_ = per_user[0, users]
        self._y_true_eval = yt[users] if len(users) != yt.shape[0] else yt
        self.__dict__.pop("_scores", None)

    @property
    def scores_(self):
        if "_scores" not in self.__dict__:
            ranks = self.y_pred_top_K_
            hits = ranks.multiply(self._y_true_eval).astype(bool).astype(np.float64).tocsr()
            hits.eliminate_zeros()
            self.__dict__["_scores"] = hits
        return self.__dict__["_scores"]

    @scores_.setter
    def scores_(self, value):
        self.__dict__["_scores"] = value

    @property
    def value(self):
        return float(self.hits_per_user_.mean()) if len(self.hits_per_user_) else float("nan")


class CoverageK(GpuTopKMixin, _bases["CoverageK"]):
    """metrics/coverage.py:13-40: fraction of all items that appear among the first K places of any evaluated
    user's list (rpk_coverage_topn)."""

    def calculate(self, y_true: csr_matrix, y_pred: csr_matrix) -> None:
        engine, idx, top_idx, top_len, yt, t_ptr_d, t_idx_d, users = self._prepare(y_true, y_pred)
        K = min(int(self.K), idx.shape[1])
        count, flags = engine.coverage_topn(yt.shape[0], idx.shape[1], K, yt.shape[1], top_idx, top_len, t_ptr_d, want_flags=True)
        self.covered_items_ = set(np.flatnonzero(flags).tolist())
        self.value_ = count / self.num_items


def _functional(cls, default_k):
    def f(y_true, y_pred, k=default_k):
        r = cls(K=k)
        r.calculate(y_true, y_pred)
        return r.value

    f.__name__ = cls.__name__.lower()
    return f


ndcg_k = _functional(NDCGK, 50)
dcg_k = _functional(DCGK, 50)
recall_k = _functional(RecallK, 50)
calibrated_recall_k = _functional(CalibratedRecallK, 50)
precision_k = _functional(PrecisionK, 10)
reciprocal_rank_k = _functional(ReciprocalRankK, 10)
hit_k = _functional(HitK, 50)
