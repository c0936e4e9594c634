"""NDCGK / DCGK / RecallK / CalibratedRecallK on the GPU -- drop-ins for recpack.metrics.

Mirror of recpack/metrics/base.py:21-295 (Metric, MetricTopK, ListwiseMetricK) and
metrics/dcg.py:21-128, metrics/recall.py:21-85: same ``calculate(y_true, y_pred)`` contract, afterwards
``value``, ``results``, ``num_users``, ``num_items``, ``name``.  Ranking of the prediction rows and
the metric itself run in rpk_topk_csr / rpk_metrics_topn; a prediction matrix produced by this
package's ``predict(..., predict_topK=N)`` carries its rank-ordered lists and skips the ranking."""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from .engine import get_engine
from .matrix import device_structure, to_host
from .util import ranks_from_lists, top_k_lists


class ListwiseMetricK:
    """metrics/base.py:253-295 (+ MetricTopK 126-193, Metric 21-123)."""

    _kind = None

    def __init__(self, K):
        self.num_users_ = 0
        self.num_items_ = 0
        self.K = K

    @property
    def name(self):
        return f"{self.__class__.__name__}_{self.K}"

    @property
    def num_items(self) -> int:
        return self.num_items_

    @property
    def num_users(self) -> int:
        return self.num_users_

    @property
    def col_names(self):
        return ["user_id", "score"]

    def _verify_shape(self, y_true, y_pred) -> bool:
        check = y_true.shape == y_pred.shape
        if not check:
            raise AssertionError(f"Shape mismatch between y_true: {y_true.shape} and y_pred: {y_pred.shape}")
        return check

    def calculate(self, y_true: csr_matrix, y_pred: csr_matrix) -> None:
        y_true = csr_matrix(y_true) if not isinstance(y_true, csr_matrix) else y_true
        self._verify_shape(y_true, y_pred)
        K = int(self.K)
        engine = get_engine()
        lists = getattr(y_pred, "_rpk_topn", None)
        top_idx = top_len = None
        if lists is not None and lists[0].shape[1] >= K and lists[0].shape[0] == y_true.shape[0]:
            idx, ln = lists
            dev = getattr(y_pred, "_rpk_topn_dev", None)
            if dev is not None and dev[2] == engine.device:
                top_idx, top_len = dev[0], dev[1]  # the lists are still on the device: nothing to upload
        else:
            idx, ln = top_k_lists(y_pred, K)
        if top_idx is None:
            top_idx, top_len = np.ascontiguousarray(idx), np.ascontiguousarray(ln)
        yt, t_ptr, t_idx, t_ptr_d, t_idx_d = device_structure(y_true, engine.device)
        U, I = yt.shape
        sums, n_users, per_user = engine.metrics_topn(U, idx.shape[1], top_idx, top_len, t_ptr_d, t_idx_d, [(self._kind, K)])
        if not isinstance(per_user, np.ndarray):
            engine.sync()
            (per_user,) = to_host(per_user)
        users = np.flatnonzero(np.diff(t_ptr) > 0)  # metrics/base.py:106-123
        self.user_id_map_ = users
        self.num_users_, self.num_items_ = len(users), I
        self.scores_ = csr_matrix(per_user[0, users].reshape(-1, 1))
        self.value_ = float(sums[0] / n_users) if n_users else float("nan")
        self._lists = (idx, ln, users, y_true.shape)

    @property
    def y_pred_top_K_(self):
        idx, ln, users, shape = self._lists
        K = int(self.K)
        return ranks_from_lists(idx[users, :K], np.minimum(ln[users], K), (len(users), shape[1]))

    @property
    def results(self):
        scores = self.scores_.toarray().ravel()
        return pd.DataFrame(dict(zip(self.col_names, (self.user_id_map_, scores))))

    @property
    def value(self):
        return self.value_


class NDCGK(ListwiseMetricK):
    """metrics/dcg.py:73-128."""

    _kind = "ndcg"


class DCGK(ListwiseMetricK):
    """metrics/dcg.py:21-52."""

    _kind = "dcg"


class RecallK(ListwiseMetricK):
    """metrics/recall.py:21-48."""

    _kind = "recall"


class CalibratedRecallK(ListwiseMetricK):
    """metrics/recall.py:58-85."""

    _kind = "calibrated_recall"


class PrecisionK(ListwiseMetricK):
    """metrics/precision.py:12-50: hits / K (fewer than K recommendations count as misses)."""

    _kind = "precision"


class ReciprocalRankK(ListwiseMetricK):
    """metrics/reciprocal_rank.py:13-40: 1 / rank of the first hit, 0 without one."""

    _kind = "reciprocal_rank"


def ndcg_k(y_true, y_pred, k=50):
    r = NDCGK(K=k)
    r.calculate(y_true, y_pred)
    return r.value


def dcg_k(y_true, y_pred, k=50):
    r = DCGK(K=k)
    r.calculate(y_true, y_pred)
    return r.value


def recall_k(y_true, y_pred, k=50):
    r = RecallK(K=k)
    r.calculate(y_true, y_pred)
    return r.value


def calibrated_recall_k(y_true, y_pred, k):
    r = CalibratedRecallK(K=k)
    r.calculate(y_true, y_pred)
    return r.value


def precision_k(y_true, y_pred, k=10):
    r = PrecisionK(K=k)
    r.calculate(y_true, y_pred)
    return r.value


def reciprocal_rank_k(y_true, y_pred, k=10):
    r = ReciprocalRankK(K=k)
    r.calculate(y_true, y_pred)
    return r.value
