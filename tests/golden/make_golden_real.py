"""Golden vectors of the real-valued fit path, made by the REAL reference (imported from /root/reference).

Run in the build container only:   python tests/golden/make_golden_real.py

Every fixture holds an input matrix, the reference's FULL item x item similarity matrix (small on purpose) and the
reference's own top-K result, so that the CUDA path can be compared with the reference's float64 values directly:
  * ItemKNN(normalize_X=True)       nearest_neighbour.py:204-224   (full = compute_* on the l1-normalised matrix)
  * compute_pearson_similarity      nearest_neighbour.py:87-111    (+ get_top_K_values, util.py:80-96)
"""
import os
import sys
import warnings

import numpy as np
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from sklearn.preprocessing import Normalizer  # noqa: E402

from recpack.algorithms import ItemKNN  # noqa: E402  (the reference)
from recpack.algorithms.nearest_neighbour import (  # noqa: E402
    compute_conditional_probability,
    compute_cosine_similarity,
    compute_pearson_similarity,
)
from recpack.util import get_top_K_values  # noqa: E402

from recpack_b200.synth import synth_interactions  # noqa: E402
from make_golden import pack  # noqa: E402


def normalize_x_case(name, X, K, similarity, pop_discount=None):
    out = {}
    X = csr_matrix(X)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo = ItemKNN(K=K, similarity=similarity, pop_discount=pop_discount, normalize_X=True)
        algo.fit(X)
        # the full matrix, exactly as ItemKNN._fit builds it (binarising wrapper base.py:129-139, then :207-215)
        Xb = X.astype(bool).astype(X.dtype)
        Xn = Normalizer(norm="l1", copy=False).transform(Xb)
        full = compute_cosine_similarity(Xn) if similarity == "cosine" else compute_conditional_probability(Xn, pop_discount)
    pack("X", X, out)
    pack("full", full, out)
    pack("S", algo.similarity_matrix_, out)
    out["K"] = np.array(K)
    out["similarity"] = np.array(similarity)
    out["pop_discount"] = np.array(np.nan if pop_discount is None else pop_discount)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "full nnz", csr_matrix(full).nnz, "S nnz", algo.similarity_matrix_.nnz)


def pearson_case(name, R, K):
    out = {}
    R = csr_matrix(R)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = compute_pearson_similarity(R)
        S = get_top_K_values(csr_matrix(full), K)
    pack("X", R, out)
    pack("full", full, out)
    pack("S", S, out)
    out["K"] = np.array(K)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "full nnz", csr_matrix(full).nnz, "S nnz", S.nnz)


def main():
    # the reference's own unit-test matrix (tests/test_algorithms/test_nearest_neighbour.py:24-32,73-107)
    data = csr_matrix(([1] * 7, ([0, 0, 1, 1, 2, 2, 2], [1, 2, 0, 2, 0, 1, 2])), shape=(4, 3))
    normalize_x_case("real_unit_normx_cosine", data, 2, "cosine")
    normalize_x_case("real_unit_normx_condprob", data, 2, "conditional_probability")
    X = synth_interactions(300, 120, 3000, seed=7)
    normalize_x_case("real_small_normx_cosine", X, 10, "cosine")
    normalize_x_case("real_small_normx_condprob", X, 10, "conditional_probability")
    normalize_x_case("real_small_normx_condprob_pd", X, 10, "conditional_probability", 0.5)
    X2 = synth_interactions(700, 300, 12000, seed=3)
    normalize_x_case("real_mid_normx_cosine", X2, 40, "cosine")

    # Pearson: the reference's unit-test matrix (test_nearest_neighbour.py:286-304) and seeded 1..5 ratings
    unit = csr_matrix(np.array([[1, 0, 1, 0], [2, 0, 2, 0], [0, 3, 0, 3], [4, 0, 0, 4]], dtype=np.float64))
    pearson_case("real_unit_pearson", unit, 2)
    rng = np.random.default_rng(5)
    R = synth_interactions(300, 120, 3000, seed=9).astype(np.float64)
    R.data[:] = rng.integers(1, 6, size=R.nnz).astype(np.float64)
    pearson_case("real_small_pearson", R, 10)


if __name__ == "__main__":
    main()
