"""Golden vectors of the TARSItemKNN family, made by the REAL reference (imported from /root/reference).

Run in the build container only:   python tests/golden/make_golden_tars.py

Per case: the interaction table (uid, iid, ts), the class name and its constructor arguments (JSON), the reference's FULL
similarity matrix of the decayed fit matrix (compute_cosine_similarity / compute_conditional_probability /
compute_pearson_similarity, time_aware_item_knn/base.py:166-181), its top-K ``similarity_matrix_`` and its ``predict``
output on the same interactions.
"""
import json
import os
import sys
import warnings

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import recpack.algorithms.time_aware_item_knn as tars  # noqa: E402
from recpack.algorithms.nearest_neighbour import (  # noqa: E402
    compute_conditional_probability,
    compute_cosine_similarity,
    compute_pearson_similarity,
)
from recpack.matrix import InteractionMatrix  # noqa: E402

from make_golden import pack  # noqa: E402

FULL = {"cosine": compute_cosine_similarity, "conditional_probability": compute_conditional_probability,
        "pearson": compute_pearson_similarity}


def case(name, cls_name, kwargs, n_users=250, n_items=90, n_rows=2500, seed=0):
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n_items + 1) ** 0.7
    df = pd.DataFrame({"uid": rng.integers(0, n_users, n_rows), "iid": rng.choice(n_items, size=n_rows, p=w / w.sum()),
                       "ts": rng.integers(0, 40 * 24 * 3600, n_rows)})
    im = InteractionMatrix(df, "iid", "uid", timestamp_ix="ts", shape=(n_users, n_items))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo = getattr(tars, cls_name)(**kwargs)
        algo.fit(im)
        pred = algo.predict(im)
        full = FULL[algo.similarity](csr_matrix(algo._add_decay_to_fit_matrix(im)))
    out = {"uid": im._df["uid"].to_numpy().astype(np.int64), "iid": im._df["iid"].to_numpy().astype(np.int64),
           "ts": im._df["ts"].to_numpy().astype(np.int64), "shape": np.array([n_users, n_items]),
           "cls": np.array(cls_name), "kwargs": np.array(json.dumps(kwargs)), "K": np.array(int(algo.K))}
    pack("full", full, out)
    pack("S", algo.similarity_matrix_, out)
    pack("pred", pred, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, cls_name, kwargs, "S nnz", algo.similarity_matrix_.nnz, "pred nnz", pred.nnz, "negative sims", int((algo.similarity_matrix_.data < 0).sum()))


if __name__ == "__main__":
    day = 1 / (24 * 3600)
    case("tars_cosine_exp", "TARSItemKNN", {"K": 10, "fit_decay": day, "predict_decay": day, "similarity": "cosine"})
    case("tars_condprob_linear", "TARSItemKNN", {"K": 10, "fit_decay": 0.5, "predict_decay": 0.3, "similarity": "conditional_probability",
                                               "decay_function": "linear", "decay_interval": 3600})
    case("tars_pearson_vaz", "TARSItemKNNVaz", {"K": 10, "fit_decay": day, "predict_decay": day / 2})
    case("tars_pearson_bigK", "TARSItemKNN", {"K": 85, "fit_decay": day, "predict_decay": day, "similarity": "pearson"})
    case("tars_liu2012", "TARSItemKNNLiu2012", {"K": 10, "decay": 2.0})
    case("tars_lee", "TARSItemKNNLee", {"K": 10, "w": 5, "similarity": "cosine"})
    case("tars_ding_nofitdecay", "TARSItemKNNDing", {"K": 10, "predict_decay": day, "similarity": "conditional_probability"})
