"""Golden vectors of FractionInteractionSplitter.split, made by the REAL reference (imported from /root/reference).

Run in the build container only:   python tests/golden/make_golden_split.py

The reference's splitter shuffles the arrays pandas hands it in place (splitters.py:247-251); with pandas 3 those are
read-only views and the unmodified call raises.  The generator therefore wraps
InteractionMatrix.interaction_history to yield copies -- the one-line fix the survey names -- and changes nothing else:
seeding, shuffle and cut are the reference's own lines.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from recpack.matrix import InteractionMatrix  # noqa: E402
from recpack.scenarios.splitters import FractionInteractionSplitter  # noqa: E402

_orig = InteractionMatrix.interaction_history.fget


def _copies(self):
    for uid, hist in _orig(self):
        yield uid, np.array(hist, copy=True)


InteractionMatrix.interaction_history = property(_copies)


def case(name, n_users, n_items, n_rows, in_frac, seed, rng_seed):
    rng = np.random.default_rng(rng_seed)
    w = 1.0 / np.arange(1, n_users + 1) ** 0.8
    uid = rng.choice(n_users, size=n_rows, p=w / w.sum())
    iid = rng.integers(0, n_items, size=n_rows)
    df = pd.DataFrame({"uid": uid, "iid": iid, "ts": rng.integers(0, 10_000, size=n_rows)})
    im = InteractionMatrix(df, "iid", "uid", timestamp_ix="ts", shape=(n_users, n_items))
    d_in, d_out = FractionInteractionSplitter(in_frac, seed=seed).split(im)
    out = {
        "uid": im._df["uid"].to_numpy().astype(np.int64),
        "iid": im._df["iid"].to_numpy().astype(np.int64),
        "interactionid": im._df["interactionid"].to_numpy().astype(np.int64),
        "in_ids": np.sort(d_in._df["interactionid"].to_numpy().astype(np.int64)),
        "out_ids": np.sort(d_out._df["interactionid"].to_numpy().astype(np.int64)),
        "in_frac": np.array(in_frac),
        "seed": np.array(seed),
        "shape": np.array([n_users, n_items]),
    }
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rows", n_rows, "in", out["in_ids"].size, "out", out["out_ids"].size, "longest history", np.bincount(uid).max())


if __name__ == "__main__":
    case("split_small", 200, 80, 4000, 0.8, 42, 1)
    case("split_heavy", 50, 500, 20000, 0.7, 7, 2)  # histories of > 1248 interactions: several generator refills
