"""EASE on the GPU (BASELINE.json configs[3]) against the reference's formulas
(recpack/algorithms/ease.py:63-95, restated in oracle.ref_ease and -- when baseline/_ref is installed -- the
reference class itself) and the reference's own unit tests (recpack/tests/test_algorithms/test_ease.py:38-87)."""
import warnings

import numpy as np
import pytest
import scipy.sparse
from scipy.sparse import csr_matrix

from conftest import HAVE_REF
from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture()
def data():
    values = [1] * 9
    users = [0, 0, 1, 1, 2, 2, 2, 3, 3]
    items = [0, 2, 0, 2, 0, 1, 2, 0, 2]
    return scipy.sparse.csr_matrix((values, (users, items)), shape=(5, 3))


def test_reference_unit_tests(data):
    """test_ease.py:38-87 (test_ease, test_alpha) through the drop-in class."""
    from recpack_b200 import EASE

    algo = EASE(l2=0.03).fit(data)
    _in = scipy.sparse.csr_matrix(([1, 1, 1], ([0, 1, 2], [0, 1, 2])), shape=(3, 3))
    result = algo.predict(_in)
    np.testing.assert_almost_equal(result[2, 0], 1, decimal=1)
    np.testing.assert_almost_equal(result[0, 2], 1, decimal=1)
    a1, a2, a3 = (EASE(l2=0.03, alpha=a).fit(data) for a in (1, 0, 2))
    np.testing.assert_almost_equal(a1.similarity_matrix_[1, 0], a2.similarity_matrix_[1, 0] / 4)
    np.testing.assert_almost_equal(a1.similarity_matrix_[2, 1], a2.similarity_matrix_[2, 1])
    np.testing.assert_almost_equal(a1.similarity_matrix_[1, 0] / 4, a3.similarity_matrix_[1, 0])


@pytest.mark.parametrize("U,I,nnz", [(300, 120, 4000), (40_000, 300, 400_000)])
def test_dense_gram_is_exact(U, I, nnz):
    """rpk_gram_dense_f64 = (X.T @ X).toarray() exactly; 40,000 users cross the 32,768-user chunk of the kernel."""
    from recpack_b200.engine import get_engine
    from recpack_b200.matrix import binary_structure
    from recpack_b200.synth import synth_interactions

    X = synth_interactions(U, I, nnz, seed=U)
    _, indptr, indices = binary_structure(X)
    G = get_engine(0).gram_dense_f64(U, I, indptr, indices)
    want = (X.T.astype(np.int64) @ X.astype(np.int64)).toarray().astype(np.float64)
    assert np.array_equal(G, want)


@pytest.mark.parametrize("alpha", [0, 0.5])
@pytest.mark.parametrize("flags", [0, 4])
def test_fit_and_predict_vs_reference_formulas(alpha, flags):
    from recpack_b200 import EASE
    from recpack_b200.engine import get_engine
    from recpack_b200.synth import synth_interactions, weak_generalization_split

    X = synth_interactions(400, 150, 6000, seed=9)
    train, _ = weak_generalization_split(X, 0.8, seed=1)
    # make sure every item is seen (alpha scaling divides by the popularity)
    train = csr_matrix(train + csr_matrix((np.ones(150, dtype=np.int32), (np.arange(150) % 400, np.arange(150))), shape=train.shape))
    train.data[:] = 1
    want_B = orc.ref_ease(train, l2=20.0, alpha=alpha).toarray()
    eng = get_engine(0)
    eng.debug_flags(flags)
    try:
        algo = EASE(l2=20.0, alpha=alpha).fit(train)
        got_B = algo.similarity_matrix_.toarray()
        np.testing.assert_allclose(got_B, want_B, rtol=1e-8, atol=1e-12)
        # scoring: with OUR model, scipy's csr @ dense on the host must give the same bits (same order of additions)
        Xb = train.astype(bool).astype(np.float64)
        host_scores = Xb @ got_B
        full = algo.predict(train).toarray()
        assert np.array_equal(full, host_scores)
        # ... and agrees with the reference's own model to rounding
        np.testing.assert_allclose(full, Xb @ want_B, rtol=1e-7, atol=1e-10)
        # top-N with the history removed: (score desc, index asc) over the non-zero scores
        N = 10
        algo.set_params(predict_topK=N, remove_history=True)
        pred = algo.predict(train)
        idx, ln = pred._rpk_topn
        masked = np.where(train.toarray() > 0, 0.0, host_scores)
        for u in range(train.shape[0]):
            cand = np.flatnonzero(masked[u])
            order = cand[np.lexsort((cand, -masked[u, cand]))][:N]
            assert ln[u] == len(order) and np.array_equal(idx[u, : ln[u]], order)
            assert np.array_equal(pred[u].toarray().ravel()[order], masked[u, order])
    finally:
        eng.debug_flags(0)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_against_the_reference_class_and_metrics():
    import recpack.algorithms
    import recpack.metrics

    from recpack_b200 import EASE, NDCGK
    from recpack_b200.synth import synth_interactions, weak_generalization_split

    X = synth_interactions(500, 200, 9000, seed=13)
    train, test_out = weak_generalization_split(X, 0.8, seed=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = recpack.algorithms.EASE(l2=50.0).fit(train)
        ours = EASE(l2=50.0).fit(train)
        assert isinstance(ours, recpack.algorithms.EASE)
        np.testing.assert_allclose(ours.similarity_matrix_.toarray(), ref.similarity_matrix_.toarray(), rtol=1e-8, atol=1e-12)
        p_ref = ref.predict(train)
        p_ref = p_ref - p_ref.multiply(train.astype(bool))
        ours.set_params(predict_topK=20, remove_history=True)
        p_ours = ours.predict(train)
    m_ref, m_ours = recpack.metrics.NDCGK(10), NDCGK(10)
    m_ref.calculate(test_out, p_ref)
    m_ours.calculate(test_out, p_ours)
    assert m_ours.value == pytest.approx(m_ref.value, rel=1e-9)


def test_ml1m_shape_topn_vs_float64_formulas():
    """Top-N exactness on the ML-1M shape (6,040 x 3,706, 1 M interactions): lists from the GPU model and scorer against
    lists ranked from the reference's formulas on the host; differences only where two scores agree to 1e-9 relative."""
    from recpack_b200 import EASE
    from recpack_b200.synth import make_dataset

    train, _, _ = make_dataset("ml1m")
    algo = EASE(l2=200.0, predict_topK=20, remove_history=True).fit(train)
    users = np.arange(0, train.shape[0], 12)
    pred = algo.predict(train[users])
    idx, ln = pred._rpk_topn
    want_B = orc.ref_ease(train, l2=200.0).toarray()
    np.testing.assert_allclose(algo.similarity_matrix_.toarray(), want_B, rtol=1e-7, atol=1e-13)
    scores = (train[users].astype(bool).astype(np.float64) @ want_B)
    scores[train[users].toarray() > 0] = 0.0
    bad = 0
    for r in range(len(users)):
        cand = np.flatnonzero(scores[r])
        order = cand[np.lexsort((cand, -scores[r, cand]))][:20]
        if not np.array_equal(idx[r, : ln[r]], order):
            # only near ties may differ: the two lists hold the same scores to 1e-9
            a, b = np.sort(scores[r, idx[r, : ln[r]]]), np.sort(scores[r, order])
            np.testing.assert_allclose(a, b, rtol=1e-9)
            bad += 1
    assert bad <= len(users) // 50
