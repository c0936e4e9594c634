#!/usr/bin/env python
"""EASE l2=200 on a named shape (default msd: 571,355 x 41,140, 33.6 M interactions): fit / predict top-20 / NDCG@10 timings
on cuda:0 and a sampled check of the scores against scipy's csr @ dense with the same model."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
from recpack_b200 import EASE, NDCGK, RecallK
from recpack_b200.engine import get_engine
from recpack_b200.synth import make_dataset

shape = sys.argv[1] if len(sys.argv) > 1 else "msd"
t0 = time.time()
train, test_out, gen = make_dataset(shape, generator="auto")
U, I = train.shape
print(f"[{shape}] data {time.time()-t0:.1f}s gen={gen} U={U} I={I} nnz={train.nnz}", flush=True)
eng = get_engine(0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        algo = EASE(l2=200.0, predict_topK=20, remove_history=True)
        algo._fit(algo._transform_fit_input(train))
        torch.cuda.synchronize(); t1 = time.time()
        pred = algo._predict(algo._transform_predict_input(train))
        torch.cuda.synchronize(); t2 = time.time()
    m = NDCGK(10); m.calculate(test_out, pred); r = RecallK(20); r.calculate(test_out, pred)
    t3 = time.time()
    print(f"[{shape}] rep{rep} fit {t1-t0:.3f} s  predict(top-20, {U} users) {t2-t1:.3f} s ({U/(t2-t1):.0f} users/s)  metrics {t3-t2:.3f} s  "
          f"NDCG@10 {m.value:.6f} Recall@20 {r.value:.6f}  scoring kernel {eng.last_timings()['predict_ms']:.1f} ms  "
          f"torch mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB", flush=True)
# sampled check: scipy's csr @ dense with the same model gives the same bits
B = algo._B_dev.cpu().numpy()
users = np.random.default_rng(0).choice(U, size=64, replace=False)
Xs = train[users].astype(bool).astype(np.float64)
want = Xs @ B
want[Xs.toarray() > 0] = 0.0
idx, ln = pred._rpk_topn
ok = 0
for r_, u in enumerate(users):
    cand = np.flatnonzero(want[r_])
    order = cand[np.lexsort((cand, -want[r_, cand]))][:20]
    ok += int(np.array_equal(idx[u, :ln[u]], order) and np.array_equal(pred[u].toarray().ravel()[order], want[r_, order]))
print(f"[{shape}] sampled users with bit-identical lists and scores vs scipy csr @ dense: {ok} of {len(users)}")
