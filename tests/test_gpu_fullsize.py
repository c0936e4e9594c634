"""Full-size parity (BASELINE.json configs[1], ML-25M shape): size-independent properties over all
rows / users, and bit-exact comparison with the oracle on samples the oracle finishes in seconds."""
import numpy as np
import pytest

from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu

K, N = 200, 20


@pytest.fixture(scope="module")
def full():
    from recpack_b200.engine import get_engine
    from recpack_b200.matrix import binary_structure
    from recpack_b200.synth import SHAPES, synth_interactions, weak_generalization_split

    U, I, nnz = SHAPES["ml25m"]
    X = synth_interactions(U, I, nnz, seed=0)
    train, test_out = weak_generalization_split(X, 0.8, seed=42)
    eng = get_engine(0)
    eng.debug_flags(0)
    _, indptr, indices = binary_structure(train)
    fit = eng.fit_topk(U, I, indptr, indices, K)
    eng.model_load_topk(I, K, fit["idx"], fit["val"], fit["len"])
    top = eng.predict_topn(U, indptr, indices, N, mask_history=True)
    return {"train": train, "test_out": test_out, "fit": fit, "top": top, "eng": eng, "indptr": indptr, "indices": indices}


def test_fit_properties_all_rows(full):
    fit, train = full["fit"], full["train"]
    I = train.shape[1]
    n = np.bincount(train.indices, minlength=I).astype(np.int64)
    idx, cnt, val, ln = fit["idx"], fit["cnt"].astype(np.int64), fit["val"], fit["len"]
    mask = np.arange(K)[None, :] < ln[:, None]
    assert np.all(idx[mask] >= 0) and np.all(idx[~mask] == -1) and np.all(cnt[~mask] == 0)
    assert not np.any(idx == np.arange(I)[:, None])  # no self similarity
    assert np.all(cnt[mask] >= 1) and np.all(cnt[mask] <= np.minimum(n[:, None], n[np.maximum(idx, 0)])[mask])
    # rank order: exact key c^2/n_j non-increasing, ties by ascending index (cross-multiplied, exact in int64 here)
    nj = n[np.maximum(idx, 0)]
    a, b = cnt[:, :-1] ** 2 * nj[:, 1:], cnt[:, 1:] ** 2 * nj[:, :-1]
    both = mask[:, 1:]
    assert np.all((a >= b)[both])
    tie = (a == b) & both
    assert np.all((idx[:, :-1] < idx[:, 1:])[tie])
    # values: c / sqrt(n_i n_j) to float64 rounding of the reference's sequential sum
    closed = cnt / np.sqrt(n[:, None].astype(np.float64) * np.maximum(nj, 1))
    np.testing.assert_allclose(val[mask], closed[mask], rtol=1e-11)
    # symmetry of the Gram: wherever i keeps j and j keeps i the counts agree
    rng = np.random.default_rng(0)
    rows = rng.choice(I, size=2000, replace=False)
    checked = 0
    for i in rows:
        for t in range(0, int(ln[i]), 37):
            j = int(idx[i, t])
            back = np.flatnonzero(idx[j, : ln[j]] == i)
            if len(back):
                assert cnt[j, back[0]] == cnt[i, t]
                checked += 1
    assert checked > 100


def test_fit_sampled_rows_bit_exact_vs_oracle(full):
    fit, train = full["fit"], full["train"]
    I = train.shape[1]
    n = np.bincount(train.indices, minlength=I)
    rng = np.random.default_rng(1)
    rows = np.unique(np.concatenate([np.argsort(-n)[:4], np.argsort(n)[:4], rng.choice(I, size=40, replace=False)]))
    want = orc.canon_fit(train, K=K, rows=rows, block=16)
    assert np.array_equal(fit["len"][rows], want["len"])
    assert np.array_equal(fit["idx"][rows], want["idx"])
    assert np.array_equal(fit["cnt"][rows], want["cnt"])
    assert np.array_equal(fit["val"][rows], want["val"])


def test_predict_sampled_users_bit_exact_vs_oracle_and_metrics(full):
    from recpack_b200.base import lists_to_csr

    fit, top, train, test_out = full["fit"], full["top"], full["train"], full["test_out"]
    U, I = train.shape
    S = lists_to_csr(fit["idx"], fit["val"], fit["len"], I)
    S.sort_indices()
    d = np.diff(train.indptr)
    rng = np.random.default_rng(2)
    users = np.unique(np.concatenate([np.argsort(-d)[:3], np.argsort(d)[:3], rng.choice(U, size=150, replace=False)]))
    want = orc.canon_predict_topn(train[users], S, N, remove_history=True)
    assert np.array_equal(top["len"][users], want["len"])
    assert np.array_equal(top["idx"][users], want["idx"])
    assert np.array_equal(top["val"][users], want["val"])
    # lists never contain history items, are sorted by (score desc, index asc)
    val, idx, ln = top["val"], top["idx"], top["len"]
    m = np.arange(N)[None, :] < ln[:, None]
    assert np.all((val[:, :-1] >= val[:, 1:])[m[:, 1:]])
    tie = (val[:, :-1] == val[:, 1:]) & m[:, 1:]
    assert np.all((idx[:, :-1] < idx[:, 1:])[tie])
    for u in users[:50]:
        assert not np.isin(idx[u, : ln[u]], train.indices[train.indptr[u] : train.indptr[u + 1]]).any()
    # metrics over all users on the GPU vs the oracle on the sample
    eng = full["eng"]
    from recpack_b200.matrix import binary_structure

    _, t_ptr, t_idx = binary_structure(test_out)
    sums, n_users, per_user = eng.metrics_topn(U, N, idx, ln, t_ptr, t_idx, [("ndcg", 10), ("recall", 20)])
    res = orc.canon_metrics_from_lists(idx[users], ln[users], test_out[users], [("ndcg", 10), ("recall", 20)])
    for m_i, key in enumerate([("ndcg", 10), ("recall", 20)]):
        value, pu, kept = res[key]
        np.testing.assert_allclose(per_user[m_i, users[kept]], pu, rtol=1e-12, atol=1e-15)
    assert n_users == int(np.count_nonzero(np.diff(test_out.indptr)))
    np.testing.assert_allclose(sums[0], np.nansum(per_user[0]), rtol=1e-10)
