"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, input coercion, constructor validation, the synthetic generator, and the multi-GPU
plumbing on the gloo backend (world size 2)."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rpk.h")).read()
    return sorted(set(re.findall(r"RPK_EXPORT\s+[\w\s\*]+?\b(rpk_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from recpack_b200 import _lib

    assert _header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    from recpack_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rpk.h")).read()
    assert lib.rpk_abi_version() == int(re.search(r"#define RPK_ABI_VERSION (\d+)", header).group(1))
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(raw, name), name


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from recpack_b200 import ItemKNN
    from recpack_b200._lib import RpkError

    with pytest.raises(RpkError):
        ItemKNN(K=2).fit(csr_matrix(np.eye(3)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "recpack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_constructor_contract():
    """recpack/algorithms/nearest_neighbour.py:170-202 and base.py:50-61."""
    from recpack_b200 import ItemKNN, NDCGK, RecallK

    a = ItemKNN(K=7, similarity="conditional_probability", pop_discount=0.5, normalize_sim=True)
    p = a.get_params()
    assert p["K"] == 7 and p["similarity"] == "conditional_probability" and p["pop_discount"] == 0.5
    assert a.identifier.startswith("ItemKNN(K=7,") and a.name == "ItemKNN" and str(a) == "ItemKNN"
    with pytest.raises(ValueError):
        ItemKNN(similarity="nope")
    with pytest.raises(ValueError):
        ItemKNN(similarity="conditional_probability", pop_discount=-0.1)
    with pytest.warns(UserWarning):
        ItemKNN(similarity="cosine", pop_discount=0.3)
    assert NDCGK(10).name == "NDCGK_10" and RecallK(20).name == "RecallK_20"
    from sklearn.exceptions import NotFittedError

    with pytest.raises(NotFittedError):
        ItemKNN().predict(csr_matrix((2, 2)))


def test_binary_structure_coercion():
    from recpack_b200.matrix import UnsupportedTypeError, binary_structure, to_csr_matrix

    X = csr_matrix((np.array([3, 1, 1, 0, 2]), np.array([2, 0, 0, 1, 1]), np.array([0, 4, 5])), shape=(2, 3))
    Xc, indptr, indices = binary_structure(X)
    assert indptr.dtype == np.int64 and indices.dtype == np.int32
    assert indptr.tolist() == [0, 2, 3] and indices.tolist() == [0, 2, 1]
    canon = csr_matrix(np.array([[1, 0, 1], [0, 1, 0]]))
    _, p2, i2 = binary_structure(canon)
    assert np.shares_memory(i2, canon.indices) or i2.tolist() == canon.indices.tolist()
    with pytest.raises(UnsupportedTypeError):
        to_csr_matrix([[1, 0]] and np.ones((2, 2)))


def test_synthetic_generator_is_exact_and_deterministic():
    from recpack_b200.synth import synth_interactions, weak_generalization_split

    X = synth_interactions(500, 200, 7000, seed=3)
    Y = synth_interactions(500, 200, 7000, seed=3)
    assert X.nnz == 7000 and X.shape == (500, 200) and X.has_canonical_format
    assert (X != Y).nnz == 0
    tr, te = weak_generalization_split(X, 0.8, seed=1)
    assert (tr + te != X).nnz == 0 and tr.multiply(te).nnz == 0
    d, dt = np.diff(X.indptr), np.diff(tr.indptr)
    assert np.array_equal(dt, np.ceil(0.8 * d).astype(dt.dtype))


def test_shard_bounds():
    from recpack_b200.distributed import shard_bounds

    cuts = shard_bounds(np.ones(10), 3)
    assert cuts[0] == 0 and cuts[-1] == 10 and all(b >= a for a, b in zip(cuts, cuts[1:]))
    cuts = shard_bounds([100, 1, 1, 1, 1, 1], 2)
    assert cuts == [0, 1, 6]
    assert shard_bounds(np.zeros(4), 2)[-1] == 4


def _gloo_worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist

    from recpack_b200.distributed import ShardExchange

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        I, K = 11, 3
        cuts = [0, 7, 11]
        rng = np.random.default_rng(5)
        full_idx = rng.integers(0, I, size=(I, K)).astype(np.int32)
        full_val = rng.random((I, K))
        full_len = rng.integers(0, K + 1, size=I).astype(np.int32)
        ex = ShardExchange(cuts, K, "cpu", dist)
        b, e = cuts[rank], cuts[rank + 1]
        a_idx, a_val, a_len = ex.gather(torch.from_numpy(full_idx[b:e]), torch.from_numpy(full_val[b:e]), torch.from_numpy(full_len[b:e]))
        assert np.array_equal(a_idx.numpy(), full_idx) and np.array_equal(a_val.numpy(), full_val)
        assert np.array_equal(a_len.numpy(), full_len)
        # packed exchange (the rows are packed by rpk_model_pack_rows on a GPU; here the send buffer is filled by hand)
        full_ent = rng.integers(0, 2**62, size=(I, K), dtype=np.int64)
        ex.p_ent[: e - b].copy_(torch.from_numpy(full_ent[b:e]))
        ex.p_len[: e - b].copy_(torch.from_numpy(full_len[b:e]))
        g_ent, g_len = ex.gather_packed()
        src = ex.row_source().numpy()
        assert np.array_equal(g_ent.numpy()[src], full_ent) and np.array_equal(g_len.numpy()[src], full_len)
        sums = torch.tensor([1.0 + rank, 2.0, 10.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(sums)
        assert sums.tolist() == [3.0, 4.0, 30.0]
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_exchange_on_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_approximate_score_margin_property():
    """The arithmetic behind k_predict_a32, emulated in numpy: a_j = sum((q >> s) | 1) with
    s = 9 + ceil(log2 d) never overflows 32 bits, stays within d of score_j / 2^s, and every item of the exact
    top N has a_j >= (N-th largest a) - 2d -- so the survivors of the approximate pass always contain it."""
    rng = np.random.default_rng(17)
    for d, n_items, N in ((1, 50, 5), (3, 200, 20), (84, 3000, 20), (700, 3000, 10), (5000, 4000, 20)):
        s = 9 + (int(np.ceil(np.log2(d))) if d > 1 else 0)
        # worst-case magnitudes included: values close to 2 (q just below 2^40) and tiny ones (q = 1)
        v = rng.random((d, n_items)) ** 4 * 1.999
        v[rng.random((d, n_items)) < 0.6] = 0.0            # K-sparse rows: most pairs absent
        q = (np.rint(v * 2.0**39).astype(np.int64) | 1) * (v > 0)
        q[0, :3] = (1 << 40) - 1                             # largest legal q
        q[-1, 3:6] = 1                                       # smallest legal q
        a_terms = np.where(q > 0, (q >> s) | 1, 0)
        score = q.sum(axis=0)
        a = a_terms.sum(axis=0)
        assert a.max() < 2**32
        cnt = (q > 0).sum(axis=0)
        assert np.all(np.abs(a - score / 2.0**s) <= cnt + 1e-9) and cnt.max() <= d
        cand = np.flatnonzero(score > 0)
        exact_top = cand[np.lexsort((cand, -score[cand]))][:N]
        a_sorted = np.sort(a[cand])[::-1]
        thr = a_sorted[min(N, len(cand)) - 1] - 2 * d
        survivors = set(cand[a[cand] >= thr].tolist())
        assert set(exact_top.tolist()) <= survivors


def test_packed_counter_carry_property():
    """The arithmetic behind the fit's packed 16-bit counters, emulated in numpy: two counts share
    a 32-bit word that is only ever incremented by 1 or 1 << 16 (in any order, also summed over several partial
    copies of the row); when the low count passes 65,535 it carries into the high half exactly (low total >> 16)
    times, so subtracting that restores the neighbour's exact count."""
    rng = np.random.default_rng(23)
    for low_total, high_total, copies in ((70_000, 1234, 1), (131_072 + 5, 65_535 - 3, 4), (65_536, 0, 3), (100, 200, 2)):
        # split both totals over `copies` partial words, each built by increments mod 2^32, then add the words
        lo_parts = rng.multinomial(low_total, np.ones(copies) / copies)
        hi_parts = rng.multinomial(high_total, np.ones(copies) / copies)
        word = 0
        for lp, hp in zip(lo_parts, hi_parts):
            part = (int(lp) + (int(hp) << 16)) & 0xFFFFFFFF      # lp increments of 1, hp increments of 1 << 16
            word = (word + part) & 0xFFFFFFFF                     # atomicAdd of the partial word
        carries = low_total >> 16
        if high_total + carries < 65_536:                         # the high half itself did not overflow
            fixed = (word - (carries << 16)) & 0xFFFFFFFF
            assert fixed >> 16 == high_total
            assert fixed & 0xFFFF == low_total & 0xFFFF


def test_registers_in_the_reference_registries():
    """the drop-in classes ARE the reference's classes (subclasses) when recpack imports, register in
    recpack's own ALGORITHM_REGISTRY / METRIC_REGISTRY (recpack/pipelines/registries.py:50-75) under new keys and
    PipelineBuilder accepts them.  Needs the reference install (baseline/install_ref.sh -> baseline/_ref)."""
    from conftest import HAVE_REF

    if not HAVE_REF:
        pytest.skip("baseline/_ref not installed")
    import recpack.algorithms
    import recpack.algorithms.base
    import recpack.metrics
    import recpack.metrics.base
    from recpack.pipelines import ALGORITHM_REGISTRY, METRIC_REGISTRY, PipelineBuilder

    import recpack_b200
    from recpack_b200 import _ref

    assert _ref.HAVE_RECPACK
    algo = recpack_b200.ItemKNN(K=5)
    assert isinstance(algo, recpack.algorithms.ItemKNN)
    assert isinstance(algo, recpack.algorithms.base.TopKItemSimilarityMatrixAlgorithm)
    assert isinstance(recpack_b200.NDCGK(10), recpack.metrics.NDCGK)
    assert isinstance(recpack_b200.RecallK(10), recpack.metrics.base.ListwiseMetricK)
    assert isinstance(recpack_b200.HitK(10), recpack.metrics.base.ElementwiseMetricK)
    assert isinstance(recpack_b200.CoverageK(10), recpack.metrics.base.GlobalMetricK)
    # the wrappers are the reference's own functions, not restatements
    assert type(algo).fit is recpack.algorithms.base.Algorithm.fit
    assert type(algo).predict is recpack.algorithms.base.Algorithm.predict

    class ItemKNNB200(recpack_b200.ItemKNN):
        pass

    class NDCGKB200(recpack_b200.NDCGK):
        pass

    if "ItemKNNB200" not in ALGORITHM_REGISTRY:
        ALGORITHM_REGISTRY.register("ItemKNNB200", ItemKNNB200)
        METRIC_REGISTRY.register("NDCGKB200", NDCGKB200)
    assert "ItemKNNB200" in ALGORITHM_REGISTRY and METRIC_REGISTRY.get("NDCGKB200") is not None
    with pytest.raises(KeyError):  # built-in names cannot be taken over (registries.py:63-75)
        ALGORITHM_REGISTRY.register("ItemKNN", ItemKNNB200)
    builder = PipelineBuilder()
    builder.add_algorithm("ItemKNNB200", params={"K": 200, "predict_topK": 20, "remove_history": True})
    builder.add_metric("NDCGKB200", K=[10])
    algo = ItemKNNB200(K=200, predict_topK=20, remove_history=True)
    assert algo.identifier.startswith("ItemKNNB200(K=200,") and algo.name == "ItemKNNB200"
    assert NDCGKB200(10).name == "NDCGKB200_10"


def test_mirror_fallback_without_recpack():
    """Without recpack (forced here with RPK_NO_RECPACK=1) the stand-alone mirror keeps the constructor contract."""
    import subprocess
    import sys

    code = (
        "import warnings, pytest\n"
        "from recpack_b200 import ItemKNN, NDCGK, HitK, CoverageK, _ref\n"
        "assert not _ref.HAVE_RECPACK\n"
        "a = ItemKNN(K=7, similarity='conditional_probability', pop_discount=0.5)\n"
        "assert a.identifier.startswith('ItemKNN(K=7,') and a.name == 'ItemKNN' and a.get_params()['pop_discount'] == 0.5\n"
        "with pytest.raises(ValueError): ItemKNN(similarity='nope')\n"
        "with pytest.raises(ValueError): ItemKNN(similarity='conditional_probability', pop_discount=1.5)\n"
        "with pytest.warns(UserWarning): ItemKNN(pop_discount=0.3)\n"
        "assert NDCGK(10).name == 'NDCGK_10' and HitK(3).name == 'HitK_3' and CoverageK(5).name == 'CoverageK_5'\n"
    )
    env = dict(os.environ, RPK_NO_RECPACK="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_only_the_binding_module_imports_recpack():
    pkg = os.path.join(ROOT, "recpack_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py") and f != "_ref.py":
            text = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+recpack(\.|\s)", text, flags=re.M), f


# ---- host helpers added in round 2 (no GPU needed)
def test_stored_zero_check_matches_numpy_all():
    from recpack_b200.matrix import _has_stored_zero

    for arr in (np.ones(10, dtype=np.int32), np.array([1, 0, 2]), np.array([-1, 2, 3]), np.array([-1, 0, 3]),
                np.array([True, True]), np.array([True, False]), np.array([0.5, 2.0]), np.array([0.5, 0.0]),
                np.array([np.nan, 1.0])):
        assert _has_stored_zero(arr) == (not bool(np.all(arr))), arr


def test_score_column_equals_scipy_route():
    from scipy.sparse import csr_matrix

    from recpack_b200.metrics import _column_csr

    for vals in (np.array([0.0, 0.25, 0.0, 1.0]), np.zeros(3), np.array([]), np.arange(1, 6, dtype=np.float64)):
        want = csr_matrix(vals.reshape(-1, 1))
        got = _column_csr(vals)
        assert got.shape == want.shape and np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data)
        assert got.mean() == want.mean() if vals.size else True


def test_splitter_validates_seed_and_handles_empty_tables_without_a_gpu():
    """np.random.RandomState(seed + u) raises for seeds outside [0, 2^32) (scenarios/splitters.py:247); so does the
    drop-in, before any device work."""
    from recpack_b200.splitters import FractionInteractionSplitter, fraction_split_mask

    with pytest.raises(ValueError):
        fraction_split_mask(np.array([0, 7, 7]), 0.5, 2**32 - 3)
    with pytest.raises(ValueError):
        fraction_split_mask(np.array([0, 1]), 0.5, -1)
    assert fraction_split_mask(np.zeros(0, dtype=np.int64), 0.5, 1).shape == (0,)
    sp = FractionInteractionSplitter(0.8, seed=42)
    assert sp.in_frac == 0.8 and sp.seed == 42 and sp.name == "FractionInteractionSplitter"


def test_interaction_matrix_fast_path_declines_without_a_gpu():
    import torch

    from recpack_b200.matrix import interaction_matrix_structure

    class Fake:
        _df = None
        shape = (2, 2)

    assert interaction_matrix_structure(Fake()) is None
    if not torch.cuda.is_available():
        import pandas as pd

        f = Fake()
        f._df = pd.DataFrame({"uid": [0, 1], "iid": [1, 0]})
        assert interaction_matrix_structure(f) is None
