"""Input coercion of the drop-in classes.

Mirrors recpack/matrix/util.py:27-77 (to_csr_matrix, UnsupportedTypeError) and the binarising
wrappers recpack/algorithms/base.py:129-151 without importing recpack: an object is accepted when it
is a scipy csr_matrix or quacks like recpack's InteractionMatrix (has ``values`` returning a CSR)."""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix


class UnsupportedTypeError(Exception):
    """Raised when a matrix of a type not supported by recpack is received (matrix/util.py:64-77)."""

    def __init__(self, X):
        super().__init__(
            "Recpack only supports matrix types InteractionMatrix, csr_matrix. Received {}.".format(type(X).__name__)
        )


def to_csr_matrix(X, binary: bool = False):
    if isinstance(X, (tuple, list)):
        return type(X)(to_csr_matrix(x, binary=binary) for x in X)
    if isinstance(X, csr_matrix):
        res = X
    elif type(X).__name__ == "InteractionMatrix" and hasattr(X, "values"):
        res = X.values
    else:
        raise UnsupportedTypeError(X)
    return binary_structure(res)[0] if binary else res


def binary_structure(X: csr_matrix):
    """(canonical CSR, indptr int64, indices int32) of the binarised matrix.

    Binarisation (recpack/util.py:99-109) only needs the sparsity structure: duplicates collapse to
    one entry and stored zeros contribute nothing, so they are dropped.  Already-canonical input
    (the normal case) is passed through without a copy."""
    if not isinstance(X, csr_matrix):
        X = csr_matrix(X)
    memo = getattr(X, "_rpk_canon", None)
    sig = (X.indptr.ctypes.data, X.indices.ctypes.data, X.data.ctypes.data, X.nnz, X.shape)
    if memo is None or memo[0] != sig:
        # validated once per matrix object (the checks are O(nnz) on the host); like scipy's own
        # has_canonical_format flag the memo assumes the arrays are not modified in place afterwards
        if X.nnz and not np.all(X.data):
            X = X.copy()
            X.eliminate_zeros()
        if not X.has_canonical_format:
            X = X.copy()
            X.sum_duplicates()
        indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(X.indices, dtype=np.int32)
        sig = (X.indptr.ctypes.data, X.indices.ctypes.data, X.data.ctypes.data, X.nnz, X.shape)
        try:
            X._rpk_canon = (sig, indptr, indices)
        except AttributeError:
            pass
        return X, indptr, indices
    return X, memo[1], memo[2]
