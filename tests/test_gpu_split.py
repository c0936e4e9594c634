"""GPU parity of the data side: FractionInteractionSplitter (rpk_split_fraction) and the device-side
InteractionMatrix -> binary CSR.  Bit-exact against the reference's goldens (tests/golden/make_golden_split.py),
against numpy's own RandomState on seeded inputs, and through the reference's InteractionMatrix when it is installed."""
import numpy as np
import pytest

from conftest import HAVE_REF, load_golden
from oracle import recpack_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["split_small", "split_heavy"])
def test_split_mask_matches_reference_golden(name):
    from recpack_b200.splitters import fraction_split_mask

    g = load_golden(name)
    mask = fraction_split_mask(g["uid"], float(g["in_frac"]), int(g["seed"]))
    assert np.array_equal(np.sort(g["interactionid"][mask]), g["in_ids"])
    assert np.array_equal(np.sort(g["interactionid"][~mask]), g["out_ids"])


@pytest.mark.parametrize("in_frac", [0.0, 0.25, 0.8, 1.0])
def test_split_mask_matches_numpy_shuffles(in_frac):
    """Histories of 1 .. 9,000 interactions (up to ~15 refills of the 624-word generator state), user ids with gaps,
    rows of a user scattered over the table."""
    from recpack_b200.splitters import fraction_split_mask

    rng = np.random.default_rng(3)
    lens = np.r_[1, 2, 3, 31, 32, 33, 623, 624, 625, 1249, 9000, rng.integers(1, 400, size=500)]
    uids = np.sort(rng.choice(100_000, size=lens.size, replace=False))
    user_ix = rng.permutation(np.repeat(uids, lens))
    seed = 12345
    got = fraction_split_mask(user_ix, in_frac, seed)
    assert np.array_equal(got, orc.ref_fraction_split_mask(user_ix, in_frac, seed))


def test_split_rejects_seeds_numpy_rejects():
    from recpack_b200.splitters import fraction_split_mask

    with pytest.raises(ValueError):
        fraction_split_mask(np.array([0, 5, 5]), 0.5, 2**32 - 3)


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference package (baseline/install_ref.sh)")
def test_splitter_dropin_on_interaction_matrix_and_device_csr():
    import pandas as pd
    from recpack.matrix import InteractionMatrix
    from recpack.scenarios.splitters import FractionInteractionSplitter as RefSplitter

    from recpack_b200.matrix import to_csr_matrix
    from recpack_b200.splitters import FractionInteractionSplitter

    g = load_golden("split_small")
    U, I = (int(v) for v in g["shape"])
    df = pd.DataFrame({"uid": g["uid"], "iid": g["iid"]})
    im = InteractionMatrix(df, "iid", "uid", shape=(U, I))
    sp = FractionInteractionSplitter(float(g["in_frac"]), seed=int(g["seed"]))
    assert isinstance(sp, RefSplitter)
    d_in, d_out = sp.split(im)
    assert np.array_equal(np.sort(d_in._df["interactionid"].to_numpy()), g["in_ids"])
    assert np.array_equal(np.sort(d_out._df["interactionid"].to_numpy()), g["out_ids"])
    assert d_in.shape == im.shape and d_out.shape == im.shape
    # the binarised matrix of the drop-in's input coercion, built on the device, equals the reference's binary_values
    for m in (im, d_in, d_out):
        got = to_csr_matrix(m, binary=True)
        want = m.binary_values
        want.sort_indices()
        assert got.shape == want.shape and np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert getattr(got, "_rpk_dev", None) is not None  # index arrays stay on the device for fit / predict
