"""EASE on the GPU -- drop-in for recpack.algorithms.EASE (recpack/algorithms/ease.py:19-109).

``_fit``: the dense Gram ``X^T X`` comes from the tensor-core kernel (exact integer counts, rpk_gram_dense_f64), the
inverse of ``X^T X + l2 I`` from cuSOLVER (Cholesky factorisation + inverse through ``torch.linalg``: the matrix is
symmetric positive definite), the closed form ``B = -P / diag(P)`` with zero diagonal and the optional popularity
scaling from rpk_ease_from_inverse.  The model stays on the device as a dense float64 matrix; ``similarity_matrix_``
(a scipy CSR like the reference's) is built on first access.

``_predict``: ``X @ B`` with the dense scorer (rpk_predict_dense_*): every score is the float64 sum over the user's
history in ascending item order -- the order of scipy's ``csr @ dense`` -- so equal models give bit-identical scores;
``predict_topK`` / ``remove_history`` as for ItemKNN.

The reference inverts with LAPACK's LU (``np.linalg.inv``); Cholesky on the GPU agrees with it to rounding
(~1e-12 relative on B), not bit for bit -- the tolerance of the tests."""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.sparse import csr_matrix
from sklearn.utils.validation import check_is_fitted

from . import _ref
from .base import lists_to_csr, matrix_signature
from .engine import get_engine
from .matrix import binary_structure, device_structure, to_csr_matrix, to_host

if _ref.HAVE_RECPACK:
    import importlib

    _EaseBase = importlib.import_module(_ref.ref_base.__name__.rsplit(".", 1)[0] + ".ease").EASE
else:
    from ._mirror import ItemSimilarityMatrixAlgorithm as _MirrorBase

    class _EaseBase(_MirrorBase):
        def __init__(self, l2=1e3, alpha=0, density=None):
            super().__init__()
            self.l2 = l2
            self.alpha = alpha
            self.density = density


class EASE(_EaseBase):
    """Embarrassingly Shallow Autoencoder (Steck 2019); arguments as in the reference (ease.py:19-62).

    :param predict_topK: optional, keep only this many scores per user in ``predict``.
    :param remove_history: optional, drop history items inside ``predict``.
    """

    def __init__(self, l2=1e3, alpha=0, density=None, predict_topK: Optional[int] = None, remove_history: bool = False):
        super().__init__(l2=l2, alpha=alpha, density=density)
        self.predict_topK = predict_topK
        self.remove_history = remove_history

    # -- fit ------------------------------------------------------------------------------
    def _transform_fit_input(self, X):
        return to_csr_matrix(X, binary=True)

    def _transform_predict_input(self, X):
        return to_csr_matrix(X, binary=True)

    def _fit(self, X: csr_matrix) -> None:
        import torch

        if self.density:
            raise NotImplementedError("EASE(density=...) (pruning of B, ease.py:97-109) is not implemented on the B200 path")
        engine = get_engine()
        X, indptr, indices, ptr_d, idx_d = device_structure(X, engine.device)
        U, I = X.shape
        G = engine.gram_dense_f64(U, I, ptr_d, idx_d)  # exact counts as float64
        engine.sync()
        n = torch.diagonal(G).clone()  # item popularities = diag(X^T X)
        # l2 * np.identity(I, dtype=float32) promotes to float64 in the sum (ease.py:80)
        torch.diagonal(G).add_(float(np.float32(self.l2)))
        L = torch.linalg.cholesky(G)       # cuSOLVER potrf
        del G
        P = torch.cholesky_inverse(L).contiguous()  # cuSOLVER potri (the result comes back column-major)
        del L
        w = None
        if self.alpha != 0:
            w = 1.0 / n.cpu().numpy() ** self.alpha  # 1 / np.diag(XTX) ** alpha (ease.py:87), inf for unseen items as there
        torch.cuda.current_stream(P.device).synchronize()
        B = engine.ease_from_inverse(P, w)  # in place
        engine.sync()
        d = self.__dict__
        d["_B_dev"] = B
        d["_similarity_host"] = None

    # -- similarity_matrix_: the dense model as the reference stores it, built on demand -------------------
    @property
    def similarity_matrix_(self):
        d = self.__dict__
        S = d.get("_similarity_host")
        if S is None:
            B = d.get("_B_dev")
            if B is None:
                raise AttributeError(f"{type(self).__name__} object has no attribute 'similarity_matrix_'")
            S = csr_matrix(B.cpu().numpy())
            d["_similarity_host"] = S
        return S

    @similarity_matrix_.setter
    def similarity_matrix_(self, S):
        d = self.__dict__
        d["_similarity_host"] = S
        d["_B_dev"] = None

    def __sklearn_is_fitted__(self):
        d = self.__dict__
        return d.get("_similarity_host") is not None or d.get("_B_dev") is not None

    def __getstate__(self):
        if self.__dict__.get("_B_dev") is not None:
            self.similarity_matrix_
        state = dict(super().__getstate__())
        state["_B_dev"] = None
        return state

    def _check_fit_complete(self):
        import warnings

        check_is_fitted(self)
        B = self.__dict__.get("_B_dev")
        if B is not None:
            missing = int((B != 0).sum(dim=1).eq(0).sum().item())
        else:
            S = csr_matrix(self.similarity_matrix_)
            missing = int(S.shape[0] - np.count_nonzero(np.asarray((S != 0).sum(axis=1)).ravel()))
        if missing > 0:
            warnings.warn(f"{self.name} missing similar items for {missing} items.")

    def _check_prediction(self, X_pred: csr_matrix, X: csr_matrix) -> None:
        import warnings

        has_hist = np.diff(X.indptr) > 0
        has_pred = np.asarray((X_pred != 0).sum(axis=1)).ravel() > 0 if X_pred.nnz and not np.all(X_pred.data) else np.diff(X_pred.indptr) > 0
        missing = int(np.count_nonzero(has_hist & ~has_pred))
        if missing > 0:
            warnings.warn(f"{self.name} failed to recommend any items for {missing} users")

    # -- predict --------------------------------------------------------------------------
    def _device_model(self, engine):
        import torch

        d = self.__dict__
        B = d.get("_B_dev")
        if B is None or B.device.index != engine.device:
            S = self.similarity_matrix_
            dense = np.ascontiguousarray(S.toarray() if hasattr(S, "toarray") else np.asarray(S), dtype=np.float64)
            B = torch.from_numpy(dense).to(torch.device("cuda", engine.device))
            torch.cuda.current_stream(B.device).synchronize()
            d["_B_dev"] = B
        return B

    def _predict(self, X: csr_matrix) -> csr_matrix:
        engine = get_engine()
        B = self._device_model(engine)
        I = int(B.shape[0])
        if X.shape[1] != I:
            raise ValueError("matmul: dimension mismatch with signature (n?,k),(k,m?)->(n?,m?)")
        U = X.shape[0]
        X, _, _, ptr_d, idx_d = device_structure(X, engine.device)
        if self.predict_topK is None:
            scores = engine.predict_dense_full(U, ptr_d, idx_d, B, mask_history=bool(self.remove_history))
            engine.sync()
            return csr_matrix(scores.cpu().numpy())
        N = int(self.predict_topK)
        top = engine.predict_dense_topn(U, ptr_d, idx_d, B, N, mask_history=bool(self.remove_history))
        engine.sync()
        idx, val, ln = to_host(top["idx"], top["val"], top["len"])
        M = lists_to_csr(idx, val, ln, I, attach=True)
        M._rpk_topn_dev = (top["idx"], top["len"], engine.device)
        M._rpk_topn_sig = matrix_signature(M)
        return M
