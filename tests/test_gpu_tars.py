"""GPU parity of the real-valued scorer (rpk_spgemm_*) and of the TARSItemKNN drop-ins.

* rpk_spgemm_*: bit-identical to scipy's ``A @ S`` (float64, csr_matmat summation order) on seeded matrices with signed
  values; top-N by (value descending, column ascending); history masking; debug-flag paths.
* TARSItemKNN family: fit compared bit for bit with the canonical top-K of the REAL reference's full similarity matrix,
  predict with scipy's product of the decayed matrix and the fitted model, and -- when no tie was broken differently --
  with the reference's own ``similarity_matrix_`` / ``predict`` output (tests/golden/make_golden_tars.py)."""
import json
import warnings

import numpy as np
import pytest
from scipy.sparse import csr_matrix, random as sprandom

from conftest import HAVE_REF, load_golden, unpack
from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu

TARS = ["tars_cosine_exp", "tars_condprob_linear", "tars_pearson_vaz", "tars_pearson_bigK", "tars_liu2012", "tars_lee",
        "tars_ding_nofitdecay"]


@pytest.fixture(scope="module")
def engine():
    from recpack_b200.engine import get_engine

    eng = get_engine(0)
    yield eng
    eng.debug_flags(0)


def _sorted(M):
    M = csr_matrix(M)
    M.sort_indices()
    return M


def _same(A, B):
    A, B = _sorted(A), _sorted(B)
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data))


def _rand_inputs(rows, I, seed, density_a=0.05, density_s=0.08):
    rs = np.random.RandomState(seed)
    A = sprandom(rows, I, density=density_a, random_state=rs, format="csr", dtype=np.float64)
    A.data[:] = np.exp(-3 * rs.rand(A.nnz))
    S = sprandom(I, I, density=density_s, random_state=rs, format="csr", dtype=np.float64)
    S.data[:] = rs.randn(S.nnz)  # signed, as Pearson similarities are
    A.sort_indices()
    S.sort_indices()
    return A, S


@pytest.mark.parametrize("flags", [0, 2, 4, 6])
@pytest.mark.parametrize("rows,I", [(1, 5), (200, 64), (700, 3000)])
def test_spgemm_csr_is_bit_identical_to_scipy(engine, rows, I, flags):
    A, S = _rand_inputs(rows, I, seed=rows + I)
    A = A.tolil()
    A[0, :] = 0  # an empty row
    A = csr_matrix(A.tocsr())
    A.eliminate_zeros()
    engine.debug_flags(flags)
    try:
        for mask in (False, True):
            indptr, indices, values = engine.spgemm_csr(A, S, mask_history=mask)
            want = csr_matrix(A @ S)
            if mask:
                want = csr_matrix(want - want.multiply(A.astype(bool)))  # pipelines/pipeline.py:174-175
            want.eliminate_zeros()
            want.sort_indices()
            assert np.array_equal(indptr, want.indptr) and np.array_equal(indices, want.indices)
            assert np.array_equal(values, want.data)  # bit-identical float64
    finally:
        engine.debug_flags(0)


@pytest.mark.parametrize("flags", [0, 2, 4])
def test_spgemm_topn_picks_by_value_then_column(engine, flags):
    A, S = _rand_inputs(300, 500, seed=9)
    S.data[:] = np.round(S.data, 1)  # ties
    N = 12
    engine.debug_flags(flags)
    try:
        got = engine.spgemm_topn(A, S, N, mask_history=True)
    finally:
        engine.debug_flags(0)
    C = csr_matrix(A @ S)
    C = csr_matrix(C - C.multiply(A.astype(bool)))
    C.eliminate_zeros()
    C.sort_indices()
    for r in range(C.shape[0]):
        lo, hi = C.indptr[r], C.indptr[r + 1]
        cols, vals = C.indices[lo:hi], C.data[lo:hi]
        order = np.lexsort((cols, -vals))[:N]
        n = order.size
        assert got["len"][r] == n
        assert np.array_equal(got["idx"][r, :n], cols[order]) and np.array_equal(got["val"][r, :n], vals[order])
        assert (got["idx"][r, n:] == -1).all()


def _lists_of(S, K):
    S = csr_matrix(S)
    idx = np.full((S.shape[0], K), -1, dtype=np.int32)
    val = np.zeros((S.shape[0], K))
    ln = np.diff(S.indptr).astype(np.int32)
    for r in range(S.shape[0]):
        lo, hi = S.indptr[r], S.indptr[r + 1]
        idx[r, : hi - lo] = S.indices[lo:hi]
        val[r, : hi - lo] = S.data[lo:hi]
    return {"idx": idx, "val": val, "len": ln}


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference package (baseline/install_ref.sh)")
@pytest.mark.parametrize("name", TARS)
def test_tars_dropins_match_the_reference(name):
    import pandas as pd
    import recpack.algorithms.time_aware_item_knn as ref_tars
    from recpack.matrix import InteractionMatrix

    import recpack_b200.time_aware as gpu_tars

    g = load_golden(name)
    U, I = (int(v) for v in g["shape"])
    df = pd.DataFrame({"uid": g["uid"], "iid": g["iid"], "ts": g["ts"]})
    im = InteractionMatrix(df, "iid", "uid", timestamp_ix="ts", shape=(U, I))
    cls_name, kwargs, K = str(g["cls"]), json.loads(str(g["kwargs"])), int(g["K"])
    algo = getattr(gpu_tars, cls_name)(**kwargs)
    assert isinstance(algo, getattr(ref_tars, cls_name))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo.fit(im)
        pred = algo.predict(im)
    # fit: the canonical top-K of the reference's own full matrix, item lists and float64 values bit for bit
    want = orc.canon_topk_of_full(unpack(g, "full"), K)
    got = _lists_of(algo.similarity_matrix_, K)
    assert np.array_equal(got["len"], want["len"]) and np.array_equal(got["idx"], want["idx"])
    assert np.array_equal(got["val"], want["val"])
    # predict: scipy's product of the reference's decayed matrix and the fitted model, bit for bit
    Xd = csr_matrix(algo._add_decay_to_predict_matrix(im))
    ref_prod = csr_matrix(Xd @ algo.similarity_matrix_)
    ref_prod.eliminate_zeros()
    assert _same(pred, ref_prod)
    # and the reference's own outputs whenever its arbitrary tie picks agree with the canonical ones
    if _same(algo.similarity_matrix_, unpack(g, "S")):
        ref_pred = unpack(g, "pred")
        ref_pred.eliminate_zeros()
        assert _same(pred, ref_pred)
    else:
        assert name in ("tars_ding_nofitdecay",)  # binary fit matrix: exact ties, the reference picks arbitrarily
