"""GPU versions of recpack/util.py:50-96 (get_top_K_ranks, get_top_K_values) with a deterministic
tie rule: value descending, then column index ascending."""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.sparse import csr_matrix

from .engine import get_engine


def top_k_lists(X: csr_matrix, K: Optional[int] = None):
    """(idx int32 [rows, K] -1 padded, len int32 [rows]) of the K best stored entries per row."""
    X = csr_matrix(X)
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    rows = X.shape[0]
    max_len = int(np.diff(X.indptr).max()) if rows else 0
    K = max(1, max_len) if K is None else int(K)
    engine = get_engine()
    idx, ln = engine.topk_csr(rows, np.ascontiguousarray(X.indptr, dtype=np.int64),
                              np.ascontiguousarray(X.indices, dtype=np.int32),
                              np.ascontiguousarray(X.data, dtype=np.float64), K)
    return idx, ln


def ranks_from_lists(idx, ln, shape) -> csr_matrix:
    rows, K = idx.shape
    mask = np.arange(K, dtype=np.int32)[None, :] < np.asarray(ln)[:, None]
    ranks = np.broadcast_to(np.arange(1, K + 1, dtype=np.int64)[None, :], idx.shape)
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(ln, out=indptr[1:])
    return csr_matrix((ranks[mask], idx[mask], indptr), shape=shape)


def get_top_K_ranks(X: csr_matrix, K: Optional[int] = None) -> csr_matrix:
    idx, ln = top_k_lists(X, K)
    return ranks_from_lists(idx, ln, X.shape)


def get_top_K_values(X: csr_matrix, K: Optional[int] = None) -> csr_matrix:
    ranks = get_top_K_ranks(X, K)
    ranks.data[:] = 1
    return ranks.multiply(X).tocsr()
