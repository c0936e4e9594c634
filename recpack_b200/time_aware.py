"""TARSItemKNN on the GPU -- drop-ins for recpack.algorithms.time_aware_item_knn.TARSItemKNN and the variants that only
fix its parameters (recpack/algorithms/time_aware_item_knn/base.py:33-201, liu_2010.py, ding_2005.py, lee_2007.py,
vaz_2013.py).

These classes exist only when ``recpack`` is importable: they need the reference's ``InteractionMatrix`` (timestamps) and
reuse its constructor validation and decay functions unchanged -- the decayed matrices are elementwise functions of the
timestamps (O(nnz) on the host, base.py:183-195).  What runs on the GPU is everything after that:

* ``_fit``: cosine / conditional-probability / Pearson similarity of the real-valued decayed matrix and its per-row top K
  (rpk_fit_topk_real: float64 sums in the reference's operation order, bit-identical values);
* ``_predict``: ``X_decayed @ similarity_matrix_`` (rpk_spgemm_count / _fill: float64, scipy's csr_matmat order, signed
  similarities allowed -- Pearson), returned as the full CSR the reference returns.

The co-occurrence-distance family (TARSItemKNNCoocDistance and its subclasses, base.py:204-330) computes a different
similarity and is not provided."""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

from . import _ref
from .engine import get_engine
from .nearest_neighbour import real_top_k

__all__ = []

if _ref.HAVE_RECPACK:
    try:
        import importlib

        _tars = importlib.import_module(_ref.ref_base.__name__.rsplit(".", 1)[0] + ".time_aware_item_knn")
    except Exception:  # pragma: no cover - environment dependent (pandas / tqdm)
        _tars = None
else:
    _tars = None


class _GpuTars:
    def _fit(self, X) -> None:
        Xd = csr_matrix(self._add_decay_to_fit_matrix(X))
        self.similarity_matrix_ = real_top_k(Xd, int(self.K), self.similarity)

    def _predict(self, X) -> csr_matrix:
        Xd = csr_matrix(self._add_decay_to_predict_matrix(X)).astype(np.float64)
        Xd.sum_duplicates()
        Xd.eliminate_zeros()
        S = self.similarity_matrix_
        S = S if isinstance(S, csr_matrix) else csr_matrix(S)
        if not S.has_sorted_indices:
            S = S.copy()
            S.sort_indices()
        if Xd.shape[1] != S.shape[0]:
            raise ValueError("matmul: dimension mismatch with signature (n?,k),(k,m?)->(n?,m?)")
        indptr, indices, values = get_engine().spgemm_csr(Xd, S)
        out = csr_matrix((values, indices, indptr), shape=(Xd.shape[0], S.shape[1]))
        out.has_canonical_format = True
        return out


if _tars is not None:
    for _name in ("TARSItemKNN", "TARSItemKNNLiu", "TARSItemKNNLiu2012", "TARSItemKNNDing", "TARSItemKNNLee", "TARSItemKNNVaz"):
        _base = getattr(_tars, _name, None)
        if _base is None:
            continue
        globals()[_name] = type(_name, (_GpuTars, _base), {"__doc__": f"GPU drop-in for recpack's {_name} (see module docstring).",
                                                          "__module__": __name__})
        __all__.append(_name)
