"""GPU parity of the real-valued fit path (rpk_fit_topk_real): ItemKNN(normalize_X=True) and Pearson.

Compared bit for bit -- item lists AND float64 values -- with the canonical top-K (value descending, item ascending,
oracle.canon_topk_of_full) of the REAL reference's full similarity matrices stored in tests/golden/real_*.npz
(tests/golden/make_golden_real.py), and on larger seeded inputs with the oracle's restatement of the same calls."""
import math
import warnings

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import load_golden, unpack
from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu

NORMX = ["real_unit_normx_cosine", "real_unit_normx_condprob", "real_small_normx_cosine", "real_small_normx_condprob",
         "real_small_normx_condprob_pd", "real_mid_normx_cosine"]
PEARSON = ["real_unit_pearson", "real_small_pearson"]
FLAG_SETS = [0, 2, 4, 6]  # bit1: tiny candidate lists; bit2: at least two column-range passes


@pytest.fixture(scope="module")
def engine():
    from recpack_b200.engine import get_engine

    eng = get_engine(0)
    yield eng
    eng.debug_flags(0)


def _params(g):
    pd_ = float(g["pop_discount"])
    return int(g["K"]), str(g["similarity"]), (None if math.isnan(pd_) else pd_)


def _lists_of(S: csr_matrix, K):
    """Rows of a CSR (stored in rank order by the drop-in) as padded lists."""
    S = csr_matrix(S)
    rows = S.shape[0]
    idx = np.full((rows, K), -1, dtype=np.int32)
    val = np.zeros((rows, K))
    ln = np.diff(S.indptr).astype(np.int32)
    for r in range(rows):
        lo, hi = S.indptr[r], S.indptr[r + 1]
        idx[r, : hi - lo] = S.indices[lo:hi]
        val[r, : hi - lo] = S.data[lo:hi]
    return {"idx": idx, "val": val, "len": ln}


def _assert_equal(got, want):
    assert np.array_equal(np.asarray(got["len"]), want["len"])
    assert np.array_equal(np.asarray(got["idx"]), want["idx"])
    assert np.array_equal(np.asarray(got["val"]), want["val"])  # bit-identical float64


def _fit_real(engine, X, values, K, sim, pd_=None, item_begin=0, item_end=None):
    X = csr_matrix(X)
    U, I = X.shape
    indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(X.indices, dtype=np.int32)
    item_pow = None
    if sim == "conditional_probability" and pd_:
        n = np.bincount(indices, minlength=I)
        item_pow = np.zeros(I)
        item_pow[n > 0] = np.power(1 / n[n > 0], pd_)
    return engine.fit_topk_real(U, I, indptr, indices, np.ascontiguousarray(values, dtype=np.float64), K, similarity=sim,
                                item_pow=item_pow, item_begin=item_begin, item_end=item_end)


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("name", NORMX)
def test_normalize_X_dropin_matches_reference_values(engine, name, flags):
    from recpack_b200 import ItemKNN

    g = load_golden(name)
    K, sim, pd_ = _params(g)
    engine.debug_flags(flags)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            algo = ItemKNN(K=K, similarity=sim, pop_discount=pd_, normalize_X=True).fit(unpack(g, "X"))
        got = _lists_of(algo.similarity_matrix_, K)
    finally:
        engine.debug_flags(0)
    _assert_equal(got, orc.canon_topk_of_full(unpack(g, "full"), K))


def test_normalize_X_reference_unit_test():
    """recpack/tests/test_algorithms/test_nearest_neighbour.py:73-107, run against the drop-in."""
    from recpack_b200 import ItemKNN

    data = csr_matrix(([1] * 7, ([0, 0, 1, 1, 2, 2, 2], [1, 2, 0, 2, 0, 1, 2])), shape=(4, 3))
    algo = ItemKNN(K=2, similarity="cosine", normalize_X=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo.fit(data)
    a, b = 1 / 9, 1 / 9 + 1 / 4
    d = math.sqrt(1 / 4 + 1 / 9)
    f = math.sqrt(2 / 4 + 1 / 9)
    expected = np.array([[0, a / (d * d), b / (d * f)], [a / (d * d), 0, b / (d * f)], [b / (d * f), b / (d * f), 0]])
    np.testing.assert_almost_equal(algo.similarity_matrix_.toarray(), expected)


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("name", PEARSON)
def test_pearson_top_k_matches_reference_values(engine, name, flags):
    from recpack_b200.nearest_neighbour import pearson_top_k

    g = load_golden(name)
    K = int(g["K"])
    engine.debug_flags(flags)
    try:
        got = _lists_of(pearson_top_k(unpack(g, "X"), K), K)
    finally:
        engine.debug_flags(0)
    _assert_equal(got, orc.canon_topk_of_full(unpack(g, "full"), K))


def test_pearson_rejects_binary_input():
    """nearest_neighbour.py:100-101, test_nearest_neighbour.py:307-311."""
    from recpack_b200.nearest_neighbour import pearson_top_k

    with pytest.raises(ValueError):
        pearson_top_k(csr_matrix(np.array([[1, 0, 1], [0, 1, 1]], dtype=np.float64)), 2)


@pytest.mark.parametrize("sim,pd_", [("cosine", None), ("conditional_probability", None), ("conditional_probability", 0.3)])
@pytest.mark.parametrize("flags", [0, 4])
def test_larger_seeded_matrix_vs_oracle_and_row_shards(engine, sim, pd_, flags):
    """2,000 x 1,500 with popular items (> 512 users: the bitmap sort of the user lists) and decayed real values, as the
    TARSItemKNN family feeds them (time_aware_item_knn/base.py:166-201); item-row shards equal the whole fit's rows."""
    from recpack_b200.synth import synth_interactions

    X = synth_interactions(2000, 1500, 60_000, seed=21)
    rng = np.random.default_rng(22)
    X = X.astype(np.float64)
    X.data[:] = np.exp(-rng.random(X.nnz) * 3.0)
    assert np.bincount(X.indices).max() > 512
    K = 25
    want = orc.canon_fit_real(X, K, sim, pd_)
    engine.debug_flags(flags)
    try:
        got = _fit_real(engine, X, X.data, K, sim, pd_)
        _assert_equal(got, want)
        b, e = 377, 1201
        part = _fit_real(engine, X, X.data, K, sim, pd_, item_begin=b, item_end=e)
        _assert_equal(part, {k: v[b:e] for k, v in want.items()})
    finally:
        engine.debug_flags(0)


def test_device_inputs_and_empty_matrix(engine):
    import torch

    from recpack_b200.synth import synth_interactions

    X = synth_interactions(400, 90, 2500, seed=4).astype(np.float64)
    X.data[:] = np.random.default_rng(1).random(X.nnz) + 0.1
    want = orc.canon_fit_real(X, 7, "cosine")
    dev = torch.device("cuda", 0)
    ptr = torch.from_numpy(X.indptr.astype(np.int64)).to(dev)
    idx = torch.from_numpy(X.indices.astype(np.int32)).to(dev)
    val = torch.from_numpy(X.data).to(dev)
    torch.cuda.synchronize()
    got = engine.fit_topk_real(400, 90, ptr, idx, val, 7)
    engine.sync()
    _assert_equal({k: v.cpu().numpy() for k, v in got.items()}, want)
    empty = engine.fit_topk_real(5, 4, np.zeros(6, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0), 3)
    assert np.array_equal(empty["len"], np.zeros(4, dtype=np.int32)) and (empty["idx"] == -1).all()
