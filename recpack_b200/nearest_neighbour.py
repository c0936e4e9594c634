"""ItemKNN on the GPU -- drop-in for recpack.algorithms.ItemKNN.

Mirror of recpack/algorithms/nearest_neighbour.py:114-224: same constructor arguments, validation
and attributes; ``_fit`` calls rpk_fit_topk instead of sklearn/scipy/numpy."""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
from scipy.sparse import csr_matrix

from .base import TopKItemSimilarityMatrixAlgorithm, lists_to_csr
from .engine import get_engine
from .matrix import device_structure, to_host


class ItemKNN(TopKItemSimilarityMatrixAlgorithm):
    """Item K Nearest Neighbours (Deshpande & Karypis 2004), cosine or conditional-probability
    similarity, K most similar items per item.  See the reference docstring
    (nearest_neighbour.py:114-167) for the model; arguments are identical.

    :param predict_topK: optional, keep only this many scores per user in ``predict``.
    :param remove_history: optional, drop history items inside ``predict``.
    """

    SUPPORTED_SIMILARITIES = ["cosine", "conditional_probability"]

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
        super().__init__(K)
        if similarity not in self.SUPPORTED_SIMILARITIES:
            raise ValueError(f"similarity {similarity} not supported")
        self.similarity = similarity
        if self.similarity != "conditional_probability" and pop_discount:
            warnings.warn(
                "Argument pop_discount is incompatible with all similarity \
                functions except conditional probability. \
                This argument will be ignored, \
                popularity discounting won't be applied.",
                UserWarning,
            )
        if type(pop_discount) == float and (pop_discount < 0 or pop_discount > 1):
            raise ValueError("Invalid value for pop_discount. Value should be between 0 and 1.")
        self.pop_discount = pop_discount
        self.normalize_X = normalize_X
        self.normalize_sim = normalize_sim
        self.predict_topK = predict_topK
        self.remove_history = remove_history

    def _fit(self, X: csr_matrix) -> None:
        if self.normalize_X:
            # l1-normalised rows make X real-valued; the GPU Gram is defined on exact integer counts
            # (SURVEY.md 8f-2).  No silent CPU path: say so.
            raise NotImplementedError("normalize_X=True is not implemented on the B200 path yet")
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
        out = engine.fit_topk(U, I, ptr_d, idx_d, K, similarity=self.similarity, item_pow=item_pow, want_cnt=False)
        if self.normalize_sim:
            # Normalizer(norm="l1") over the kept entries of each row (nearest_neighbour.py:220-222), on the host
            idx, val, ln = to_host(out["idx"], out["val"], out["len"])
            mask = np.arange(idx.shape[1])[None, :] < ln[:, None]
            row_sum = np.where(mask, np.abs(val), 0.0).sum(axis=1)
            row_sum[row_sum == 0] = 1.0
            self.similarity_matrix_ = lists_to_csr(idx, np.ascontiguousarray(val / row_sum[:, None]), ln, I)
        else:
            self._set_device_fit(out, I, engine.device)
