"""Algorithm base classes with the reference's fit / predict contract.

Mirror of recpack/algorithms/base.py:33-304 (Algorithm, ItemSimilarityMatrixAlgorithm,
TopKItemSimilarityMatrixAlgorithm): same names, same wrappers, same log line and warnings.  The
scoring of ItemSimilarityMatrixAlgorithm runs on the GPU (rpk_predict_*), there is no CPU path."""
from __future__ import annotations

import logging
import time
import warnings

import numpy as np
from scipy.sparse import csr_matrix
from sklearn.base import BaseEstimator
from sklearn.utils.validation import check_is_fitted

from .engine import get_engine
from .matrix import binary_structure, device_structure, to_csr_matrix, to_host

logger = logging.getLogger("recpack")


class Algorithm(BaseEstimator):
    """recpack/algorithms/base.py:33-217."""

    def __init__(self):
        super().__init__()

    @property
    def name(self):
        return self.__class__.__name__

    @property
    def identifier(self):
        paramstring = ",".join((f"{k}={v}" for k, v in self.get_params().items()))
        return self.name + "(" + paramstring + ")"

    def __str__(self):
        return self.name

    def set_params(self, **params):
        super().set_params(**params)

    def _fit(self, X: csr_matrix):
        raise NotImplementedError("Please implement _fit")

    def _predict(self, X: csr_matrix) -> csr_matrix:
        raise NotImplementedError("Please implement _predict")

    def _check_fit_complete(self):
        check_is_fitted(self)

    def _check_prediction(self, X_pred: csr_matrix, X: csr_matrix) -> None:
        """Warn when a user with history got no recommendation (base.py:108-127); computed from the
        row pointers instead of Python sets over nonzero()."""
        has_hist = np.diff(X.indptr) > 0
        has_pred = np.diff(X_pred.indptr) > 0
        if X_pred.nnz and not np.all(X_pred.data):  # explicit zeros do not count as recommendations
            has_pred = np.asarray((X_pred != 0).sum(axis=1)).ravel() > 0
        missing = int(np.count_nonzero(has_hist & ~has_pred))
        if missing > 0:
            warnings.warn(f"{self.name} failed to recommend any items for {missing} users")

    def _transform_fit_input(self, X):
        return to_csr_matrix(X, binary=True)

    def _transform_predict_input(self, X):
        return to_csr_matrix(X, binary=True)

    def fit(self, X):
        start = time.time()
        X = self._transform_fit_input(X)
        self._fit(X)
        self._check_fit_complete()
        end = time.time()
        logger.info(f"Fitting {self.name} complete - Took {end - start :.3}s")
        return self

    def predict(self, X) -> csr_matrix:
        self._check_fit_complete()
        X = self._transform_predict_input(X)
        X_pred = self._predict(X)
        self._check_prediction(X_pred, X)
        return X_pred


class ItemSimilarityMatrixAlgorithm(Algorithm):
    """recpack/algorithms/base.py:220-279: predict = X @ similarity_matrix_, on the GPU.

    Two additions follow the reference's own precedent ``predict_topK`` ("Use when the user x item
    output matrix would become too large for RAM", base.py:427-430):

    * ``predict_topK``: keep only the N best scores per user (score desc, item index asc);
    * ``remove_history``: drop the user's own history items inside predict -- before the truncation,
      which is where the pipeline removes them (pipelines/pipeline.py:174-175).

    With both unset the full score matrix of the reference is returned."""

    predict_topK = None
    remove_history = False

    # -- similarity_matrix_: host CSR, materialised on first use when the fit result lives on the device ----
    @property
    def similarity_matrix_(self):
        d = self.__dict__
        S = d.get("_similarity_host")
        if S is None:
            if d.get("_fit_dev") is None:
                raise AttributeError(f"{type(self).__name__} object has no attribute 'similarity_matrix_'")
            S = self._materialize_similarity()
        return S

    @similarity_matrix_.setter
    def similarity_matrix_(self, S):
        self.__dict__["_similarity_host"] = S
        self.__dict__["_fit_dev"] = None  # an assigned matrix replaces the device-resident fit result

    def _materialize_similarity(self):
        dev = self.__dict__["_fit_dev"]
        idx, val, ln = to_host(dev["idx"], dev["val"], dev["len"])
        S = lists_to_csr(idx, val, ln, dev["I"])
        self.__dict__["_similarity_host"] = S
        return S

    def _set_device_fit(self, out, I, device):
        """Keep the rank-ordered top-K lists of a fit on the device (torch tensors owned by this object)."""
        d = self.__dict__
        get_engine(device).sync()  # the lists are complete before torch touches them (streams may differ)
        d["_similarity_host"] = None
        d["_fit_dev"] = {"idx": out["idx"], "val": out["val"], "len": out["len"], "I": int(I), "device": int(device),
                         "empty_rows": int((out["len"] == 0).sum().item())}

    def __sklearn_is_fitted__(self):
        d = self.__dict__
        return d.get("_similarity_host") is not None or d.get("_fit_dev") is not None

    def __getstate__(self):
        if self.__dict__.get("_fit_dev") is not None:
            self.similarity_matrix_  # pickles carry the host matrix, never device memory
        state = dict(super().__getstate__())
        state["_fit_dev"] = None
        return state

    # -- device model management ---------------------------------------------------------------
    def _device_model_key(self):
        S = self.similarity_matrix_
        return (id(S), S.shape, S.nnz, S.data.ctypes.data, S.indices.ctypes.data)

    def _ensure_device_model(self, engine):
        dev = self.__dict__.get("_fit_dev")
        if dev is not None and dev["device"] == engine.device:
            key = ("dev", engine.device, id(dev["idx"]))
            if getattr(engine, "_model_key", None) != key:
                engine.model_load_topk(dev["I"], dev["idx"].shape[1], dev["idx"], dev["val"], dev["len"])
                engine._model_key = key
                engine._model_owner = dev["idx"]  # keeps the id unique while it is the loaded model
            return
        S = self.similarity_matrix_
        if not isinstance(S, csr_matrix):
            S = csr_matrix(S)
            self.similarity_matrix_ = S
        key = (engine.device,) + self._device_model_key()
        if getattr(engine, "_model_key", None) == key:
            return
        if not S.has_canonical_format:
            S = S.copy()
            S.sum_duplicates()
        if S.nnz and not np.all(S.data):
            S = S.copy()
            S.eliminate_zeros()
        engine.model_load_csr(S.shape[0], np.ascontiguousarray(S.indptr, dtype=np.int64),
                              np.ascontiguousarray(S.indices, dtype=np.int32),
                              np.ascontiguousarray(S.data, dtype=np.float64))
        engine._model_key = key
        engine._model_owner = self.__dict__.get("_similarity_host")

    def _n_items(self):
        dev = self.__dict__.get("_fit_dev")
        return dev["I"] if dev is not None else self.similarity_matrix_.shape[0]

    def _predict(self, X: csr_matrix) -> csr_matrix:
        engine = get_engine()
        I = self._n_items()
        if X.shape[1] != I:
            raise ValueError("matmul: dimension mismatch with signature (n?,k),(k,m?)->(n?,m?)")
        self._ensure_device_model(engine)
        U = X.shape[0]
        if self.predict_topK is None:
            X, indptr, indices = binary_structure(X)
            o_ptr, o_idx, o_val = engine.predict_csr(U, indptr, indices, mask_history=bool(self.remove_history))
            return csr_matrix((o_val, o_idx, o_ptr), shape=(U, I))
        N = int(self.predict_topK)
        X, _, _, ptr_d, idx_d = device_structure(X, engine.device)
        top = engine.predict_topn(U, ptr_d, idx_d, N, mask_history=bool(self.remove_history))
        engine.sync()
        idx, val, ln = to_host(top["idx"], top["val"], top["len"])
        M = lists_to_csr(idx, val, ln, I, attach=True)
        M._rpk_topn_dev = (top["idx"], top["len"], engine.device)  # the metrics read the lists where they are
        return M

    def _check_fit_complete(self):
        super()._check_fit_complete()
        assert self.__sklearn_is_fitted__()  # hasattr(self, "similarity_matrix_") without building the matrix
        dev = self.__dict__.get("_fit_dev")
        if dev is not None:
            missing = dev["empty_rows"]
        else:
            S = csr_matrix(self.similarity_matrix_)
            rows_with_score = np.diff(S.indptr) > 0
            if S.nnz and not np.all(S.data):
                rows_with_score = np.asarray((S != 0).sum(axis=1)).ravel() > 0
            missing = int(S.shape[0] - np.count_nonzero(rows_with_score))
        if missing > 0:
            warnings.warn(f"{self.name} missing similar items for {missing} items.")


class TopKItemSimilarityMatrixAlgorithm(ItemSimilarityMatrixAlgorithm):
    """recpack/algorithms/base.py:282-304."""

    def __init__(self, K):
        super().__init__()
        self.K = K


def lists_to_csr(idx, val, ln, n_cols, attach=False) -> csr_matrix:
    """[rows x K] rank-ordered lists (-1 padded) -> CSR whose rows keep the rank order.

    When every row is full the arrays are used as they are (no copy)."""
    rows, K = idx.shape
    ln = np.asarray(ln)
    if rows and int(ln.min()) == K:
        indptr = np.arange(rows + 1, dtype=np.int64) * K
        M = csr_matrix((val.reshape(-1), idx.reshape(-1), indptr), shape=(rows, n_cols))
    else:
        mask = np.arange(K, dtype=np.int32)[None, :] < ln[:, None]
        indptr = np.zeros(rows + 1, dtype=np.int64)
        np.cumsum(ln, out=indptr[1:])
        M = csr_matrix((val[mask], idx[mask], indptr), shape=(rows, n_cols))
    if attach:
        M._rpk_topn = (idx, ln)
    return M
