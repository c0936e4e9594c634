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


# ----------------------------------------------------------------------------------------------
# The same generator and split on a CUDA device (torch): the 100 M / 500 M interaction shapes take
# minutes with numpy on the host and seconds here.  Harness only (torch is plumbing); the data differ
# from the numpy generator's for the same seed, so a run names the generator it used.
# ----------------------------------------------------------------------------------------------
def synth_interactions_cuda(U: int, I: int, nnz: int, seed: int = 0, item_exp: float = 0.9, user_exp: float = 0.6,
                            device=None) -> csr_matrix:
    import torch

    if nnz > U * I:
        raise ValueError("more interactions than cells")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    pu = torch.arange(1, U + 1, dtype=torch.float64, device=dev) ** -user_exp
    pi = torch.arange(1, I + 1, dtype=torch.float64, device=dev) ** -item_exp
    cu = torch.cumsum(pu / pu.sum(), 0)
    ci = torch.cumsum(pi / pi.sum(), 0)
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    chunk = 1 << 27  # candidate pairs drawn per round (bounds the temporaries)
    while keys.numel() < nnz:
        want = min(int((nnz - keys.numel()) * 1.5) + 1024, chunk)
        u = torch.searchsorted(cu, torch.rand(want, dtype=torch.float64, device=dev, generator=g)).clamp_(max=U - 1)
        i = torch.searchsorted(ci, torch.rand(want, dtype=torch.float64, device=dev, generator=g)).clamp_(max=I - 1)
        keys = torch.unique(torch.cat([keys, u * I + i]))  # sorted, duplicates removed
        del u, i
    if keys.numel() > nnz:
        drop = torch.randperm(keys.numel(), device=dev, generator=g)[: keys.numel() - nnz]
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=dev)
        keep[drop] = False
        keys = keys[keep]
        del drop, keep
    perm_u = torch.randperm(U, device=dev, generator=g)
    perm_i = torch.randperm(I, device=dev, generator=g)
    keys = perm_u[keys // I] * I + perm_i[keys % I]
    keys = torch.sort(keys).values
    out = _csr_from_sorted_keys(keys.cpu().numpy(), U, I)
    del keys
    torch.cuda.empty_cache()
    return out


def weak_generalization_split_cuda(X: csr_matrix, frac_in: float = 0.8, seed: int = 42, device=None):
    """weak_generalization_split on a CUDA device (same semantics, its own random stream)."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    U, I = X.shape
    indptr = torch.from_numpy(np.ascontiguousarray(X.indptr, dtype=np.int64)).to(dev)
    indices = torch.from_numpy(np.ascontiguousarray(X.indices, dtype=np.int64)).to(dev)
    d = indptr[1:] - indptr[:-1]
    rows = torch.repeat_interleave(torch.arange(U, dtype=torch.int64, device=dev), d)
    # random order inside each user: sort by (user, random 31-bit key)
    rk = torch.randint(0, 1 << 31, (X.nnz,), dtype=torch.int64, device=dev, generator=g)
    order = torch.argsort(rows * (1 << 31) + rk)
    del rk
    pos = torch.arange(X.nnz, dtype=torch.int64, device=dev) - indptr[:-1][rows]
    n_in = torch.ceil(frac_in * d.to(torch.float64)).to(torch.int64)
    to_in = pos < n_in[rows]
    keys = rows * I + indices[order]
    k_in = torch.sort(keys[to_in]).values.cpu().numpy()
    k_out = torch.sort(keys[~to_in]).values.cpu().numpy()
    del keys, order, pos, rows
    torch.cuda.empty_cache()
    return _csr_from_sorted_keys(k_in, U, I), _csr_from_sorted_keys(k_out, U, I)


def make_dataset(shape: str, seed: int = 0, split_seed: int = 42, generator: str = "auto"):
    """(train, test_out, generator name) of a named shape.  generator: "numpy" (the reference generator of
    SURVEY.md 8d), "cuda", or "auto" = numpy up to 25 M interactions, cuda above when a device is there."""
    U, I, nnz = SHAPES[shape]
    if generator == "auto":
        generator = "numpy"
        if nnz > 30_000_000:
            try:
                import torch

                if torch.cuda.is_available():
                    generator = "cuda"
            except ImportError:
                pass
    if generator == "cuda":
        X = synth_interactions_cuda(U, I, nnz, seed=seed)
        train, test_out = weak_generalization_split_cuda(X, 0.8, seed=split_seed)
    else:
        X = synth_interactions(U, I, nnz, seed=seed)
        train, test_out = weak_generalization_split(X, 0.8, seed=split_seed)
    return train, test_out, generator
