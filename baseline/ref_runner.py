"""Row-blocked runner of the REFERENCE's own CPU implementation of the hot path (bench.py's reference arm and
`cpu_baseline` leg).  Harness code: nothing here is imported by the product.

The unmodified reference (installed by baseline/install_ref.sh into baseline/_ref) cannot run the ML-25M shape in
one piece: its Gram has 2.5e9 stored entries and `get_top_K_values` dies allocating two more copies (SURVEY.md
0.3).  BASELINE.md section 3 therefore times the SAME calls on row blocks, which is bit-identical per row:

  fit, item rows `blk`   sklearn `cosine_similarity(Xt[blk], Xt, dense_output=False)`     (nearest_neighbour.py:80)
                         or `invert(diag(n[blk])) @ (Xb.T[blk] @ X)`                      (nearest_neighbour.py:48-61)
                         -> diagonal zeroed in place                                      (:64, :81)
                         -> recpack.util.get_top_K_values(S_blk, K)                       (:216, util.py:80-96)
  scoring, users `blk`   recpack ItemKNN._predict(X[blk])  (= X @ similarity_matrix_)     (algorithms/base.py:237-255)
                         -> Algorithm._check_prediction                                   (base.py:108-127)
                         -> X_pred - X_pred.multiply(X_in)                                (pipelines/pipeline.py:174-175)
                         -> recpack.metrics.NDCGK(10) / RecallK(20).calculate             (metrics/base.py:172-193)

When `recpack` cannot be imported the oracle's restatement of the same calls is used and the result says
kind = "port"."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def import_reference():
    """The reference's modules from baseline/_ref (None when it is not installed)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "recpack")):
        return None
    for p in (os.path.join(ROOT, "baseline", "stubs"), ref_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import recpack.algorithms
        import recpack.algorithms.util
        import recpack.metrics
        import recpack.util

        return {"ItemKNN": recpack.algorithms.ItemKNN, "get_top_K_values": recpack.util.get_top_K_values,
                "invert": recpack.algorithms.util.invert, "NDCGK": recpack.metrics.NDCGK, "RecallK": recpack.metrics.RecallK,
                "file": recpack.__file__}
    except Exception:  # pragma: no cover - environment dependent
        return None


class RefRunner:
    """Holds the matrices the blocks need (built once, outside every timed region)."""

    def __init__(self, train: csr_matrix, test_out: csr_matrix, K: int, similarity: str):
        self.ref = import_reference()
        self.kind = "reference" if self.ref is not None else "port"
        self.train, self.test_out, self.K, self.similarity = train, test_out, int(K), similarity
        X = csr_matrix(train)
        self.Xb = X.astype(bool).astype(X.dtype)  # the binarising wrapper, base.py:129-139 / util.py:99-109
        if similarity == "cosine":
            self.Xt = self.Xb.astype(np.float64).T.tocsr()  # what cosine_similarity(X.T) works on
        else:
            self.Xt = self.Xb.T.tocsr()
            self.n = np.asarray(self.Xb.sum(axis=0)).ravel()
        self.algo = None

    # ---- fit of a block of item rows -----------------------------------------------------------
    def fit_rows(self, rows) -> csr_matrix:
        rows = np.asarray(rows)
        if self.similarity == "cosine":
            from sklearn.metrics.pairwise import cosine_similarity

            Sb = csr_matrix(cosine_similarity(self.Xt[rows], self.Xt, dense_output=False))
        else:
            co = csr_matrix(self.Xt[rows] @ self.Xb)  # to_binary(X).T @ X, the block's rows
            if self.ref is not None:
                A = self.ref["invert"](sp.diags(self.n[rows]).tocsr())
            else:
                inv = np.zeros(len(rows))
                nz = self.n[rows] > 0
                inv[nz] = 1 / self.n[rows][nz]
                A = sp.diags(inv).tocsr()
            Sb = csr_matrix(A @ co)
        # setdiag(0) of the full matrix overwrites the stored diagonal in place; the same here
        for r, i in enumerate(rows):
            lo, hi = Sb.indptr[r], Sb.indptr[r + 1]
            hit = np.flatnonzero(Sb.indices[lo:hi] == i)
            if len(hit):
                Sb.data[lo + hit[0]] = 0
        if self.ref is not None:
            return csr_matrix(self.ref["get_top_K_values"](Sb, K=self.K))
        from oracle import recpack_oracle as orc

        return csr_matrix(orc.ref_top_k_values(Sb, self.K))

    def fit_all(self, procs: int, block: int = 1024) -> csr_matrix:
        """The complete similarity matrix, item-row blocks dealt to `procs` forked workers (setup for the scoring
        leg, not timed as part of a step: the reference itself is single-threaded)."""
        I = self.Xt.shape[0]
        blocks = [np.arange(s, min(I, s + block)) for s in range(0, I, block)]
        global _RUNNER
        _RUNNER = self
        if procs <= 1:
            parts = [self.fit_rows(b) for b in blocks]
        else:
            import multiprocessing as mp

            with mp.get_context("fork").Pool(procs) as pool:
                parts = pool.map(_fit_block_worker, blocks, chunksize=1)
        return sp.vstack(parts).tocsr()

    # ---- scoring of a block of users ---------------------------------------------------------------
    def set_model(self, S: csr_matrix):
        self.S = csr_matrix(S)
        if self.ref is not None:
            self.algo = self.ref["ItemKNN"](K=self.K, similarity=self.similarity)
            self.algo.similarity_matrix_ = self.S

    def score_users(self, users):
        """predict -> check -> history removal -> NDCG@10, Recall@20.  Returns (sum ndcg, sum recall, users counted)."""
        import warnings

        users = np.asarray(users)
        Xin, Yout = self.Xb[users], self.test_out[users]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if self.algo is not None:
                pred = self.algo._predict(Xin)
                self.algo._check_prediction(pred, Xin)
                pred = pred - pred.multiply(Xin)
                m1, m2 = self.ref["NDCGK"](10), self.ref["RecallK"](20)
                m1.calculate(Yout, pred)
                m2.calculate(Yout, pred)
                n = m1.num_users
                return float(m1.value) * n, float(m2.value) * n, int(n)
            from oracle import recpack_oracle as orc

            pred = orc.ref_remove_history(orc.ref_predict(Xin, self.S), Xin)
            v1, pu1, _ = orc.ref_ndcg(Yout, pred, 10)
            v2, pu2, _ = orc.ref_recall(Yout, pred, 20)
            return float(np.sum(pu1)), float(np.sum(pu2)), int(len(pu1))

    # ---- one bounded step: the same fraction of the item rows and of the users ---------------------
    def step(self, fraction: float, seed: int):
        rng = np.random.default_rng(seed)
        U, I = self.train.shape
        rows = np.sort(rng.choice(I, size=max(1, int(round(fraction * I))), replace=False))
        users = np.sort(rng.choice(U, size=max(1, int(round(fraction * U))), replace=False))
        t0 = time.perf_counter()
        for s in range(0, len(rows), 2048):
            self.fit_rows(rows[s : s + 2048])
        t1 = time.perf_counter()
        nd, rc, n = 0.0, 0.0, 0
        for s in range(0, len(users), 8192):
            a, b, c = self.score_users(users[s : s + 8192])
            nd, rc, n = nd + a, rc + b, n + c
        t2 = time.perf_counter()
        return {"seconds": t2 - t0, "fit_seconds": t1 - t0, "score_seconds": t2 - t1, "rows": len(rows), "users": len(users),
                "ndcg10": nd / max(n, 1), "recall20": rc / max(n, 1)}


_RUNNER = None


def _fit_block_worker(rows):
    os.environ["OMP_NUM_THREADS"] = "1"
    return _RUNNER.fit_rows(rows)


def _step_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    fraction, seed = args
    return _RUNNER.step(fraction, seed)


def parallel_steps(runner: RefRunner, fraction: float, seeds):
    """One step per forked worker, all at once: what the box's host cores deliver together on independent blocks."""
    import multiprocessing as mp

    global _RUNNER
    _RUNNER = runner
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(seeds)) as pool:
        outs = pool.map(_step_worker, [(fraction, s) for s in seeds], chunksize=1)
    return outs, time.perf_counter() - t0
