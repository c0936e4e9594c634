"""The tensor-core Gram (tcgen05 int8 MMA, int32 TMEM accumulation) against a dense numpy product."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(256, 128), (300, 200), (129, 33), (1000, 1024), (2100, 777), (4000, 4096)])
def test_dense_gram_is_exact(shape):
    from recpack_b200.engine import get_engine

    I, Kd = shape
    rng = np.random.default_rng(I * 7 + Kd)
    dens = rng.random(Kd) ** 2  # columns of very different density
    A = (rng.random((I, Kd)) < dens[None, :]).astype(np.uint8)
    G = get_engine(0).gram_dense_u16(A)
    want = (A.astype(np.float32) @ A.astype(np.float32).T).astype(np.int64)  # exact: counts < 2^24
    assert G.dtype == np.uint16 and G.shape == (I, I)
    assert np.array_equal(G.astype(np.int64), want)


def test_dense_gram_all_ones_and_zeros():
    from recpack_b200.engine import get_engine

    eng = get_engine(0)
    A = np.ones((384, 640), dtype=np.uint8)
    assert np.all(eng.gram_dense_u16(A) == 640)
    A[:] = 0
    assert np.all(eng.gram_dense_u16(A) == 0)
