#!/usr/bin/env python
"""Times the GPU FractionInteractionSplitter mask on an ML-25M-shape interaction table (25 M rows) and checks a sample
of users against numpy's RandomState.  usage: probe_split.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recpack_b200.splitters import fraction_split_mask
from recpack_b200.synth import synth_interactions_cuda

X = synth_interactions_cuda(162541, 59047, 25_000_095, seed=0)
user_ix = np.repeat(np.arange(X.shape[0], dtype=np.int64), np.diff(X.indptr))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    mask = fraction_split_mask(user_ix, 0.8, 42)
    t1 = time.time()
    print(f"[split ml25m shape, {user_ix.size} rows] rep{rep} {1e3*(t1-t0):.1f} ms (upload + stable sort + rpk_split_fraction + mask back), in = {mask.sum()}", flush=True)
rng = np.random.default_rng(0)
bad = 0
for u in rng.choice(X.shape[0], size=300, replace=False):
    b, e = X.indptr[u], X.indptr[u + 1]
    h = np.arange(b, e); np.random.RandomState(42 + int(u)).shuffle(h)
    want = np.zeros(e - b, bool); want[h[: int(np.ceil((e - b) * 0.8))] - b] = True
    bad += not np.array_equal(mask[b:e], want)
print("sampled users differing from numpy:", bad)
t0 = time.time()
for u in range(2000):
    b, e = X.indptr[u], X.indptr[u + 1]
    h = np.arange(b, e); np.random.RandomState(42 + u).shuffle(h)
print(f"numpy loop (the reference's inner loop, without pandas): {(time.time()-t0)/2000*X.shape[0]:.1f} s extrapolated for all users")
