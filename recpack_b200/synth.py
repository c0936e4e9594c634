"""Synthetic interaction data of the shapes named in BASELINE.json.

Harness code for bench.py and the tests: a seeded power-law generator and a
WeakGeneralization-style split.  Nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

# name -> (users, items, interactions)
SHAPES = {
    "ml100k": (943, 1_682, 100_000),
    "ml1m": (6_040, 3_706, 1_000_000),
    "ml25m": (162_541, 59_047, 25_000_000),
    "netflix": (480_189, 17_770, 100_000_000),
    "msd": (571_355, 41_140, 33_600_000),
    "large": (1_000_000, 200_000, 500_000_000),
}


def synth_interactions(U: int, I: int, nnz: int, seed: int = 0, item_exp: float = 0.9, user_exp: float = 0.6) -> csr_matrix:
    """Binary CSR (users x items, int32 data, sorted int32 indices) with exactly ``nnz``
    unique pairs.  Item weight ~ rank^-item_exp, user weight ~ rank^-user_exp, pairs drawn
    i.i.d. from the product, de-duplicated and topped up; ids are randomly permuted because
    RecPack's id spaces are first-appearance order, not popularity order
    (recpack/preprocessing/preprocessors.py:205-219)."""
    if nnz > U * I:
        raise ValueError("more interactions than cells")
    rng = np.random.default_rng(seed)
    pu = np.arange(1, U + 1, dtype=np.float64) ** -user_exp
    pi = np.arange(1, I + 1, dtype=np.float64) ** -item_exp
    cu = np.cumsum(pu / pu.sum())
    ci = np.cumsum(pi / pi.sum())
    keys = np.empty(0, dtype=np.int64)
    while len(keys) < nnz:
        want = int((nnz - len(keys)) * 1.5) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(want)), U - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(ci, rng.random(want)), I - 1).astype(np.int64)
        fresh = u * I + i
        fresh.sort()
        fresh = fresh[np.r_[True, fresh[1:] != fresh[:-1]]]
        if len(keys):
            fresh = fresh[~np.isin(fresh, keys, assume_unique=True)]
            keys = np.concatenate([keys, fresh])
            keys.sort()
        else:
            keys = fresh
    if len(keys) > nnz:
        keep = np.zeros(len(keys), dtype=bool)
        keep[rng.choice(len(keys), size=nnz, replace=False)] = True
        keys = keys[keep]
    perm_u = rng.permutation(U).astype(np.int64)
    perm_i = rng.permutation(I).astype(np.int64)
    keys = perm_u[keys // I] * I + perm_i[keys % I]
    keys.sort()
    return _csr_from_sorted_keys(keys, U, I)


def _csr_from_sorted_keys(keys: np.ndarray, U: int, I: int) -> csr_matrix:
    """Binary CSR from sorted unique keys u * I + i."""
    u = keys // I
    indptr = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(np.bincount(u, minlength=U), out=indptr[1:])
    ptr_dtype = np.int32 if len(keys) < 2**31 else np.int64
    X = csr_matrix((np.ones(len(keys), dtype=np.int32), (keys % I).astype(np.int32), indptr.astype(ptr_dtype)), shape=(U, I))
    X.has_sorted_indices = True
    return X


def weak_generalization_split(X: csr_matrix, frac_in: float = 0.8, seed: int = 42):
    """Per user, a random ``ceil(frac_in * d_u)`` of the interactions go to train (= the
    fold-in history at test time), the rest to test_out -- the semantics of
    recpack/scenarios/weak_generalization.py:105-121 with splitters.py:247-256.
    Returns (train, test_out) as binary CSR with sorted indices."""
    X = csr_matrix(X)
    rng = np.random.default_rng(seed)
    U, I = X.shape
    d = np.diff(X.indptr).astype(np.int64)
    rows = np.repeat(np.arange(U, dtype=np.int64), d)
    order = np.argsort(rows + rng.random(X.nnz))  # random order inside each user, users stay grouped
    pos = np.arange(X.nnz, dtype=np.int64) - np.repeat(X.indptr[:-1].astype(np.int64), d)
    to_in = pos < np.repeat(np.ceil(frac_in * d).astype(np.int64), d)
    keys = rows * I + X.indices[order].astype(np.int64)  # rows[order] == rows: users stay grouped
    k_in = keys[to_in]
    k_in.sort()
    k_out = keys[~to_in]
    k_out.sort()
    return _csr_from_sorted_keys(k_in, U, I), _csr_from_sorted_keys(k_out, U, I)
