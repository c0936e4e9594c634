"""FractionInteractionSplitter on the GPU -- drop-in for recpack.scenarios.splitters.FractionInteractionSplitter
(recpack/scenarios/splitters.py:212-263), the per-user random split behind WeakGeneralization / StrongGeneralization.

The reference loops over the users in Python: ``np.random.RandomState(seed + u).shuffle(history)``, the first
``ceil(n * in_frac)`` interaction ids go to ``data_in`` -- minutes at ML-25M size, and more than fit + predict + metrics
together once those run on the GPU.  Here the grouping by user is a stable device sort (torch: plumbing)
and the shuffles run in ``rpk_split_fraction``, which reproduces numpy's MT19937 stream and Fisher-Yates shuffle bit for
bit, so both halves hold exactly the interactions the reference would put there."""
from __future__ import annotations

import numpy as np

from .engine import get_engine

from . import _ref

if _ref.HAVE_RECPACK and _ref.ref_splitters is not None:
    # subclass the reference's splitter when recpack is importable (same registry / isinstance behaviour)
    _Base = _ref.ref_splitters.FractionInteractionSplitter
else:  # stand-alone mirror of the constructor (splitters.py:222-231)

    class _Base:  # type: ignore
        def __init__(self, in_frac, seed: int = None):
            self.in_frac = in_frac
            if seed is None:
                seed = np.random.get_state()[1][0]
            self.seed = seed

        @property
        def name(self):
            return self.__class__.__name__

        @property
        def identifier(self):
            paramstring = ",".join((f"{k}={v}" for k, v in self.__dict__.items()))
            return self.name + f"({paramstring})"


USER_IX = "uid"  # InteractionMatrix.USER_IX (matrix/interaction_matrix.py:50-53)


def fraction_split_mask(user_ix: np.ndarray, in_frac: float, seed: int, device: int = 0) -> np.ndarray:
    """Boolean mask over the interaction table's rows: True = ``data_in``.  ``user_ix``: the user of every row."""
    import torch

    user_ix = np.ascontiguousarray(user_ix, dtype=np.int64)
    n_rows = user_ix.shape[0]
    if n_rows == 0:
        return np.zeros(0, dtype=bool)
    lo, hi = int(user_ix.min()), int(user_ix.max())
    if int(seed) + lo < 0 or int(seed) + hi > 2**32 - 1:
        raise ValueError("Seed must be between 0 and 2**32 - 1")  # numpy's own error for RandomState(seed + u)
    dev = torch.device("cuda", device)
    u = torch.from_numpy(user_ix).to(dev)
    u_sorted, rows = torch.sort(u, stable=True)  # rows of a user keep the table order (pandas groupby does too)
    uids, counts = torch.unique_consecutive(u_sorted, return_counts=True)
    seg = torch.zeros(uids.shape[0] + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=seg[1:])
    torch.cuda.current_stream(dev).synchronize()  # the library runs on its own stream
    engine = get_engine(device)
    mask = engine.split_fraction(uids.contiguous(), seg, rows.contiguous(), float(in_frac), int(seed))
    engine.sync()
    return mask.cpu().numpy().astype(bool)


class FractionInteractionSplitter(_Base):
    """Split data randomly, such that ``in_frac`` of every user's interactions go to the first return value and the
    remainder to the second; arguments and results as the reference's class."""

    def split(self, data):
        mask = fraction_split_mask(data._df[USER_IX].to_numpy(), self.in_frac, int(self.seed), get_engine().device)
        return data._apply_mask(mask), data._apply_mask(~mask)
