"""ItemKNN on the GPU -- drop-in for recpack.algorithms.ItemKNN.

With ``recpack`` importable this class subclasses the reference's own ``ItemKNN``
(recpack/algorithms/nearest_neighbour.py:114-224): constructor validation, warnings and attributes are inherited,
``_fit`` calls rpk_fit_topk instead of sklearn / scipy / numpy, ``_predict`` is the GPU scorer of
``GpuSimilarityMixin``."""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.sparse import csr_matrix

from . import _ref
from .base import GpuSimilarityMixin, lists_to_csr
from .engine import get_engine
from .matrix import device_structure, to_host

if _ref.HAVE_RECPACK:
    _ItemKNNBase = _ref.ref_nn.ItemKNN
else:
    from ._mirror import ItemKNNArgs as _ItemKNNBase


class ItemKNN(GpuSimilarityMixin, _ItemKNNBase):
    """Item K Nearest Neighbours (Deshpande & Karypis 2004), cosine or conditional-probability
    similarity, K most similar items per item.  See the reference docstring
    (nearest_neighbour.py:114-167) for the model; arguments are identical.

    :param predict_topK: optional, keep only this many scores per user in ``predict``.
    :param remove_history: optional, drop history items inside ``predict``.
    """

    def __init__(
        self,
        K=200,
        similarity: str = "cosine",
        pop_discount: Optional[float] = None,
        normalize_X: bool = False,
        normalize_sim: bool = False,
        predict_topK: Optional[int] = None,
        remove_history: bool = False,
    ):
        super().__init__(K=K, similarity=similarity, pop_discount=pop_discount, normalize_X=normalize_X,
                         normalize_sim=normalize_sim)
        self.predict_topK = predict_topK
        self.remove_history = remove_history

    def _fit(self, X: csr_matrix) -> None:
        engine = get_engine()
        X, indptr, indices, ptr_d, idx_d = device_structure(X, engine.device)
        U, I = X.shape
        item_pow = None
        if self.similarity == "conditional_probability" and self.pop_discount:
            # A.power(pop_discount) of the reference (nearest_neighbour.py:58), same numpy call
            n = np.bincount(indices, minlength=I)
            item_pow = np.zeros(I, dtype=np.float64)
            nz = n > 0
            item_pow[nz] = np.power(1 / n[nz], self.pop_discount)
        K = int(self.K)
        # the rank-ordered lists stay on the device (torch tensors owned by this estimator); similarity_matrix_
        # is built from them on first use
        out = engine.fit_topk(U, I, ptr_d, idx_d, K, similarity=self.similarity, item_pow=item_pow, want_cnt=False,
                              normalize_X=bool(self.normalize_X))
        if self.normalize_sim:
            # Normalizer(norm="l1") over the kept entries of each row (nearest_neighbour.py:220-222), on the host
            idx, val, ln = to_host(out["idx"], out["val"], out["len"])
            mask = np.arange(idx.shape[1])[None, :] < ln[:, None]
            row_sum = np.where(mask, np.abs(val), 0.0).sum(axis=1)
            row_sum[row_sum == 0] = 1.0
            self.similarity_matrix_ = lists_to_csr(idx, np.ascontiguousarray(val / row_sum[:, None]), ln, I)
        else:
            self._set_device_fit(out, I, engine.device)
