"""GPU tests of the drop-in boundary: the classes driven by the REFERENCE's own wrappers, registries and
``Pipeline.run()`` (recpack/pipelines/pipeline.py:135-179), the two warnings of the wrappers, and the metrics added
from the same top-N lists (HitK, CoverageK).  The reference is the unmodified install under baseline/_ref
(baseline/install_ref.sh); tests that need it are skipped when it is absent."""
import warnings

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import HAVE_REF
from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed (baseline/install_ref.sh)")


def _data(U=400, I=150, nnz=6000, seed=3):
    from recpack_b200.synth import synth_interactions, weak_generalization_split

    X = synth_interactions(U, I, nnz, seed=seed)
    return weak_generalization_split(X, 0.8, seed=seed + 1)


def test_failed_to_recommend_warning():
    """a8: recpack/algorithms/base.py:108-127 -- users WITH history that get no recommendation are counted in a
    warning; users without history are not."""
    from recpack_b200 import ItemKNN

    # items 0,1 co-occur (users 0,1); item 2 is only ever seen alone (users 2,3) -> no neighbours -> users 2 and 3
    # have history but cannot be recommended anything; user 4 has no history at all
    X = csr_matrix(np.array([[1, 1, 0], [1, 1, 0], [0, 0, 1], [0, 0, 1], [0, 0, 0]], dtype=np.int32))
    with pytest.warns(UserWarning, match="ItemKNN missing similar items for 1 items."):
        algo = ItemKNN(K=2).fit(X)
    for kw in ({}, {"predict_topK": 2}, {"predict_topK": 2, "remove_history": True}):
        algo.set_params(**kw)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            pred = algo.predict(X)
        msgs = [str(x.message) for x in w]
        # remove_history empties the lists of users 0 and 1 as well (their only candidates are their own items)
        expect = 4 if kw.get("remove_history") else 2
        assert f"ItemKNN failed to recommend any items for {expect} users" in msgs, msgs
        assert "ItemKNN missing similar items for 1 items." in msgs  # predict re-runs the fit check (base.py:209)
        assert pred.shape == X.shape and pred[4].nnz == 0


@needs_ref
def test_failed_to_recommend_warning_matches_reference():
    import recpack.algorithms

    from recpack_b200 import ItemKNN

    X = csr_matrix(np.array([[1, 1, 0], [1, 1, 0], [0, 0, 1], [0, 0, 1], [0, 0, 0]], dtype=np.int32))

    def run(cls):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            cls(K=2).fit(X).predict(X)
        return sorted(str(x.message) for x in w if issubclass(x.category, UserWarning))

    assert run(ItemKNN) == run(recpack.algorithms.ItemKNN)


def test_hitk_and_coveragek_from_lists():
    """metrics/hit.py:20-45, metrics/coverage.py:13-40 against a direct numpy evaluation of the same lists."""
    from recpack_b200 import CoverageK, HitK, ItemKNN

    train, test_out = _data()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pred = ItemKNN(K=20, predict_topK=20, remove_history=True).fit(train).predict(train)
    idx, ln = pred._rpk_topn
    users = np.flatnonzero(np.diff(test_out.indptr) > 0)
    for K in (1, 5, 20):
        h = HitK(K)
        h.calculate(test_out, pred)
        c = CoverageK(K)
        c.calculate(test_out, pred)
        hits, covered = [], set()
        for u in users:
            top = idx[u, : min(K, ln[u])]
            covered.update(top.tolist())
            hits.append(np.isin(top, test_out.indices[test_out.indptr[u] : test_out.indptr[u + 1]]).sum())
        assert h.value == pytest.approx(np.mean(hits), abs=1e-15)
        assert h.num_users == len(users) and h.name == f"HitK_{K}"
        assert np.array_equal(np.asarray(h.scores_.sum(axis=1)).ravel(), np.array(hits, dtype=np.float64))
        assert c.value == len(covered) / train.shape[1] and c.covered_items_ == covered
        res = h.results
        assert list(res.columns) == ["user_id", "item_id", "score"] and res["score"].sum() == np.sum(hits)


@needs_ref
@pytest.mark.parametrize("K", [3, 10])
def test_all_metrics_match_reference_classes_on_the_same_prediction(K):
    """Every metric class of this package against the reference's class of the same name, both fed the SAME
    prediction matrix whose scores have no ties (so the reference's arbitrary tie picks cannot differ)."""
    import recpack.metrics as rm

    import recpack_b200 as rb

    rng = np.random.default_rng(7)
    U, I = 300, 90
    dense = rng.random((U, I)) * (rng.random((U, I)) < 0.3)
    dense[:5] = 0  # users without any prediction
    y_pred = csr_matrix(dense)
    y_true = csr_matrix((rng.random((U, I)) < 0.05).astype(np.int32))
    for name in ("NDCGK", "DCGK", "RecallK", "CalibratedRecallK", "PrecisionK", "ReciprocalRankK", "HitK", "CoverageK"):
        ours, ref = getattr(rb, name)(K), getattr(rm, name)(K)
        ours.calculate(y_true, y_pred)
        ref.calculate(y_true, y_pred)
        assert ours.name == ref.name and ours.num_users == ref.num_users and ours.num_items == ref.num_items
        assert ours.value == pytest.approx(ref.value, rel=1e-12, abs=1e-15), name
        a, b = ours.results, ref.results
        assert list(a.columns) == list(b.columns) and len(a) == len(b), name
        if "item_id" in a.columns:
            key = ["user_id", "item_id"]
            a, b = a.sort_values(key).reset_index(drop=True), b.sort_values(key).reset_index(drop=True)
            assert np.array_equal(a["user_id"], b["user_id"])
            assert np.array_equal(a["item_id"], b["item_id"], equal_nan=True)  # users without recommendations: NaN rows
        elif "user_id" in a.columns:
            a, b = a.sort_values("user_id").reset_index(drop=True), b.sort_values("user_id").reset_index(drop=True)
            assert np.array_equal(a["user_id"], b["user_id"])
        np.testing.assert_allclose(a["score"].to_numpy(), b["score"].to_numpy(), rtol=1e-12, atol=1e-15, err_msg=name)


def test_edited_prediction_matrix_is_ranked_again():
    """A prediction matrix edited in place after predict() no longer matches the lists attached to it: the metric
    must rank the matrix as it is now."""
    from recpack_b200 import ItemKNN, NDCGK

    train, test_out = _data(seed=11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pred = ItemKNN(K=20, predict_topK=10).fit(train).predict(train)
    m0 = NDCGK(10)
    m0.calculate(test_out, pred)
    plain = csr_matrix((pred.data.copy(), pred.indices.copy(), pred.indptr.copy()), shape=pred.shape)
    # zero the best item of every user in place
    first = pred.indptr[:-1][np.diff(pred.indptr) > 0]
    pred.data[first] = 0.0
    plain.data[first] = 0.0
    m1, m2 = NDCGK(10), NDCGK(10)
    m1.calculate(test_out, pred)
    m2.calculate(test_out, plain)  # no attachments at all
    assert m1.value == m2.value and m1.value != m0.value


@needs_ref
@pytest.mark.parametrize("similarity,K", [("cosine", 20), ("conditional_probability", 10)])
def test_reference_pipeline_runs_the_dropin(similarity, K):
    """recpack/pipelines/pipeline.py:135-179: the reference's own Pipeline.run() -- fit(InteractionMatrix) ->
    predict(InteractionMatrix) -> X_pred - X_pred.multiply(history) -> metric.calculate -- with the drop-in
    algorithm and metrics registered next to the built-in ones.  Compared with the built-in ItemKNN in the same
    pipeline (tie picks of the reference aside) and bit for bit with the canonical oracle."""
    import pandas as pd
    from recpack.matrix import InteractionMatrix
    from recpack.pipelines import ALGORITHM_REGISTRY, METRIC_REGISTRY, PipelineBuilder

    import recpack_b200

    class ItemKNNB200(recpack_b200.ItemKNN):
        pass

    class NDCGKB200(recpack_b200.NDCGK):
        pass

    class RecallKB200(recpack_b200.RecallK):
        pass

    for key, cls, reg in (("ItemKNNB200", ItemKNNB200, ALGORITHM_REGISTRY), ("NDCGKB200", NDCGKB200, METRIC_REGISTRY),
                          ("RecallKB200", RecallKB200, METRIC_REGISTRY)):
        if key not in reg:
            reg.register(key, cls)

    train, test_out = _data(U=500, I=200, nnz=9000, seed=21)

    def to_im(M):
        coo = M.tocoo()
        df = pd.DataFrame({"uid": coo.row, "iid": coo.col, "ts": np.arange(coo.nnz)})
        return InteractionMatrix(df, "iid", "uid", timestamp_ix="ts", shape=M.shape)

    im_train, im_out = to_im(train), to_im(test_out)
    builder = PipelineBuilder("b200-dropin-test")
    builder.set_full_training_data(im_train)
    builder.set_test_data((im_train, im_out))
    builder.add_algorithm("ItemKNN", params={"K": K, "similarity": similarity})
    builder.add_algorithm("ItemKNNB200", params={"K": K, "similarity": similarity})
    builder.add_algorithm("ItemKNNB200", params={"K": K, "similarity": similarity, "predict_topK": 20, "remove_history": True})
    for name in ("NDCGK", "NDCGKB200", "RecallKB200"):
        builder.add_metric(name, K=[10])
    pipeline = builder.build()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pipeline.run()
    res = pipeline.get_metrics()
    assert len(res) == 3
    rows = {k.split("(")[0] + ("+topn" if "predict_topK=20" in k else ""): v for k, v in res.to_dict("index").items()}
    ref_row, full_row, topn_row = rows["ItemKNN"], rows["ItemKNNB200"], rows["ItemKNNB200+topn"]
    # the drop-in algorithm under the reference's metric and under its own metric: same prediction matrix; they agree
    # to rounding unless scores tie exactly (conditional probability: sums of c/n_i), where the reference's ranking
    # picks arbitrarily among equals (recpack/util.py:68) and ours takes the smaller index
    assert full_row["NDCGK_10"] == pytest.approx(full_row["NDCGKB200_10"], rel=1e-12 if similarity == "cosine" else 2e-2)
    # full-CSR predict and fused top-N predict feed the same lists to the metrics
    assert topn_row["NDCGKB200_10"] == pytest.approx(full_row["NDCGKB200_10"], rel=1e-12)
    assert topn_row["RecallKB200_10"] == pytest.approx(full_row["RecallKB200_10"], rel=1e-12)
    # against the built-in ItemKNN: the only differences are the reference's arbitrary tie picks
    assert full_row["NDCGKB200_10"] == pytest.approx(ref_row["NDCGK_10"], rel=2e-2)
    # bit-exact target: the canonical oracle on the same data
    want = orc.canon_fit(train, K=K, similarity=similarity)
    S = orc.topk_to_csr(want["idx"], want["val"], want["len"], train.shape[1])
    top = orc.canon_predict_topn(train, S, 10, remove_history=True)
    val = orc.canon_metrics_from_lists(top["idx"], top["len"], test_out, [("ndcg", 10), ("recall", 10)])
    assert topn_row["NDCGKB200_10"] == pytest.approx(val[("ndcg", 10)][0], rel=1e-12)
    assert topn_row["RecallKB200_10"] == pytest.approx(val[("recall", 10)][0], rel=1e-12)


def test_postfilters_inside_predict_refill_the_lists():
    """postprocessing/filters.py:58-101 applied inside predict: equal to filtering the UNTRUNCATED prediction matrix and
    ranking afterwards -- the places of the removed items are taken by the next best ones."""
    from recpack_b200 import ExcludeItems, ItemKNN, SelectItems
    from recpack_b200.util import top_k_lists

    train, _ = _data(U=300, I=120, nnz=5000, seed=31)
    N = 10
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo = ItemKNN(K=30).fit(train)
        full = algo.predict(train)  # every score, no filter
        popular = np.argsort(-np.bincount(train.indices, minlength=120))[:15]
        keep = np.arange(0, 120, 2)
        for filters in ([ExcludeItems(popular)], [SelectItems(keep)], [ExcludeItems(popular), SelectItems(keep)]):
            want = full
            for f in filters:
                want = csr_matrix(f.apply(want))
            want.eliminate_zeros()
            want = csr_matrix(want - want.multiply(train.astype(bool)))  # history removal (pipeline.py:174-175)
            want.eliminate_zeros()
            w_idx, w_len = top_k_lists(want, N)
            algo.set_params(predict_topK=N, remove_history=True)
            got = algo.set_postfilters(filters).predict(train)
            g_idx, g_len = got._rpk_topn
            assert np.array_equal(g_len, w_len) and np.array_equal(g_idx, w_idx)
            allowed = np.ones(120, dtype=bool)
            for f in filters:
                allowed &= f.item_mask(120).astype(bool)
            assert allowed[g_idx[g_idx >= 0]].all()
            # the full-matrix mode honours the filters too
            algo.set_params(predict_topK=None, remove_history=False)
            dense = algo.predict(train).toarray()
            assert not dense[:, ~allowed].any()
            ref_dense = full.toarray()
            np.testing.assert_array_equal(dense[:, allowed], ref_dense[:, allowed])
        algo.set_postfilters(None)
        algo.set_params(predict_topK=None, remove_history=False)
        assert (algo.predict(train) != full).nnz == 0  # the filter is gone again


def test_truncated_model_equals_a_fit_at_the_smaller_K():
    """A sweep over K fits once: the first K places of the rank-ordered lists are the fit at K, bit for bit."""
    from recpack_b200 import ItemKNN

    train, _ = _data(U=500, I=200, nnz=9000, seed=41)
    for sim in ("cosine", "conditional_probability"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            big = ItemKNN(K=60, similarity=sim, predict_topK=10, remove_history=True).fit(train)
            for K in (1, 7, 60):
                small = ItemKNN(K=K, similarity=sim, predict_topK=10, remove_history=True).fit(train)
                cut = big.truncated(K)
                assert cut.K == K and cut.get_params()["similarity"] == sim
                a, b = cut.similarity_matrix_.copy(), small.similarity_matrix_.copy()
                a.sort_indices()
                b.sort_indices()
                assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
                pa, pb = cut.predict(train), small.predict(train)
                assert np.array_equal(pa._rpk_topn[0], pb._rpk_topn[0]) and np.array_equal(pa.data, pb.data)
