"""Full-size parity at the sizes BASELINE.json names: size-independent properties over all rows / users,
and bit-exact comparison with the oracle on samples the oracle finishes in seconds.

  configs[1]  ItemKNN cosine K=200, ML-25M shape (162,541 x 59,047, 25 M interactions)
  configs[2]  ItemKNN conditional probability K=100, Netflix shape (480,189 x 17,770, 100 M interactions):
              123 items are seen by >= 65,536 users (more than HEAVY_CAP = 32), so those rows take the 32-bit
              counter launch of the fit; integer keys tie in most rows
  configs[4]  ItemKNN cosine K=100, 1,000,000 x 200,000 with 500 M interactions: several item-range passes in
              the fit (P = 2) and in the scoring kernels, dense leg off
"""
import numpy as np
import pytest

from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu

N = 20

CONFIGS = {
    "ml25m_cosine": dict(shape="ml25m", similarity="cosine", K=200, generator="numpy"),
    "netflix_condprob": dict(shape="netflix", similarity="conditional_probability", K=100, generator="cuda"),
    "large_cosine": dict(shape="large", similarity="cosine", K=100, generator="cuda"),
}


def _run_config(name):
    import torch

    from recpack_b200.engine import get_engine
    from recpack_b200.matrix import binary_structure
    from recpack_b200.synth import make_dataset

    cfg = CONFIGS[name]
    train, test_out, _ = make_dataset(cfg["shape"], seed=0, split_seed=42, generator=cfg["generator"])
    U, I = train.shape
    K = cfg["K"]
    eng = get_engine(0)
    eng.debug_flags(0)
    _, indptr, indices = binary_structure(train)
    ptr_d = torch.from_numpy(indptr).cuda()
    idx_d = torch.from_numpy(indices).cuda()
    fit_d = eng.fit_topk(U, I, ptr_d, idx_d, K, similarity=cfg["similarity"])
    eng.model_load_topk(I, K, fit_d["idx"], fit_d["val"], fit_d["len"])
    top_d = eng.predict_topn(U, ptr_d, idx_d, N, mask_history=True)
    eng.sync()
    fit = {k: v.cpu().numpy() for k, v in fit_d.items()}
    top = {k: v.cpu().numpy() for k, v in top_d.items()}
    del fit_d, top_d, ptr_d, idx_d
    torch.cuda.empty_cache()
    return {"train": train, "test_out": test_out, "fit": fit, "top": top, "eng": eng, "cfg": cfg, "prep": orc.prepare(train)}


@pytest.fixture(scope="module", params=list(CONFIGS))
def full(request):
    out = _run_config(request.param)
    yield out
    out.clear()


def test_fit_properties_all_rows(full):
    fit, train, cfg = full["fit"], full["train"], full["cfg"]
    K, cosine = cfg["K"], cfg["similarity"] == "cosine"
    I = train.shape[1]
    n = np.bincount(train.indices, minlength=I).astype(np.int64)
    idx, cnt, val, ln = fit["idx"], fit["cnt"].astype(np.int64), fit["val"], fit["len"]
    mask = np.arange(K)[None, :] < ln[:, None]
    assert np.all(idx[mask] >= 0) and np.all(idx[~mask] == -1) and np.all(cnt[~mask] == 0)
    assert not np.any(idx == np.arange(I)[:, None])  # no self similarity
    nj = n[np.maximum(idx, 0)]
    assert np.all(cnt[mask] >= 1) and np.all(cnt[mask] <= np.minimum(n[:, None], nj)[mask])
    both = mask[:, 1:]
    if cosine:
        # rank order: exact key c^2/n_j non-increasing, ties by ascending index (cross-multiplied; c^2 * n < 2^63 here)
        a, b = cnt[:, :-1] ** 2 * nj[:, 1:], cnt[:, 1:] ** 2 * nj[:, :-1]
    else:
        a, b = cnt[:, :-1], cnt[:, 1:]  # conditional probability without pop_discount: integer key c
    assert np.all((a >= b)[both])
    tie = (a == b) & both
    assert np.all((idx[:, :-1] < idx[:, 1:])[tie])
    if cosine:
        # values: c / sqrt(n_i n_j) to float64 rounding of the reference's sequential sum of c equal terms
        # (each of the c additions rounds once: relative error <= c * 2^-53)
        closed = cnt / np.sqrt(n[:, None].astype(np.float64) * np.maximum(nj, 1))
        np.testing.assert_allclose(val[mask], closed[mask], rtol=max(1e-11, 2.0 * float(cnt.max()) * 2.0**-53))
    else:
        # fl(fl(1/n_i) * c): bit-exact closed form (algorithms/util.py:132, nearest_neighbour.py:53)
        inv = np.zeros(I, dtype=np.float64)
        inv[n > 0] = 1.0 / n[n > 0]
        assert np.array_equal(val[mask], (inv[:, None] * cnt)[mask])
    # a full row cannot have left out anything it co-occurs with more often: rows shorter than K hold every neighbour
    short = np.flatnonzero((ln < K) & (n > 0))[:50]
    if len(short):
        co = orc.cooccurrence_rows(full["prep"], short)
        for r, i in enumerate(short):
            c_row = np.asarray(co[r]).ravel().copy()
            c_row[i] = 0
            assert ln[i] == np.count_nonzero(c_row)
    # symmetry of the Gram: wherever i keeps j and j keeps i the counts agree
    rng = np.random.default_rng(0)
    rows = rng.choice(I, size=min(I, 6000), replace=False)
    checked = 0
    for i in rows:
        for t in range(0, int(ln[i]), 17):
            j = int(idx[i, t])
            back = np.flatnonzero(idx[j, : ln[j]] == i)
            if len(back):
                assert cnt[j, back[0]] == cnt[i, t]
                checked += 1
    assert checked > 100


def _tie_heavy_rows(fit, want_rows):
    """Rows whose K-th and (K+1)-th keys are most likely tied: the kept tail has equal counts."""
    cnt, ln = fit["cnt"], fit["len"]
    K = cnt.shape[1]
    full_rows = np.flatnonzero(ln == K)
    tail_tied = full_rows[cnt[full_rows, K - 1] == cnt[full_rows, K - 2]]
    return tail_tied[:want_rows]


def test_fit_sampled_rows_bit_exact_vs_oracle(full):
    fit, train, cfg = full["fit"], full["train"], full["cfg"]
    I = train.shape[1]
    n = np.bincount(train.indices, minlength=I)
    rng = np.random.default_rng(1)
    heavy = np.argsort(-n)[:4]
    rows = np.unique(np.concatenate([heavy, np.argsort(n)[:4], rng.choice(I, size=40, replace=False), _tie_heavy_rows(fit, 40),
                                     np.flatnonzero(n >= 65536)[:6]]))
    want = orc.canon_fit(full["prep"], K=cfg["K"], similarity=cfg["similarity"], rows=rows, block=16)
    assert np.array_equal(fit["len"][rows], want["len"])
    assert np.array_equal(fit["idx"][rows], want["idx"])
    assert np.array_equal(fit["cnt"][rows], want["cnt"])
    assert np.array_equal(fit["val"][rows], want["val"])


def test_predict_sampled_users_bit_exact_vs_oracle_and_metrics(full):
    from recpack_b200.base import lists_to_csr
    from recpack_b200.matrix import binary_structure

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
    _, t_ptr, t_idx = binary_structure(test_out)
    sums, n_users, per_user = eng.metrics_topn(U, N, idx, ln, t_ptr, t_idx, [("ndcg", 10), ("recall", 20)])
    res = orc.canon_metrics_from_lists(idx[users], ln[users], test_out[users], [("ndcg", 10), ("recall", 20)])
    for m_i, key in enumerate([("ndcg", 10), ("recall", 20)]):
        value, pu, kept = res[key]
        np.testing.assert_allclose(per_user[m_i, users[kept]], pu, rtol=1e-12, atol=1e-15)
    assert n_users == int(np.count_nonzero(np.diff(test_out.indptr)))
    np.testing.assert_allclose(sums[0], np.nansum(per_user[0]), rtol=1e-10)
