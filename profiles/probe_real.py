#!/usr/bin/env python
"""Times ItemKNN(normalize_X=True) (rpk_fit_topk_real) on a named shape on cuda:0 with the phase trace.
usage: probe_real.py <shape> <similarity> <K>"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recpack_b200.engine import get_engine
from recpack_b200.synth import make_dataset

shape, sim, K = sys.argv[1], sys.argv[2], int(sys.argv[3])
train, test_out, gen = make_dataset(shape, generator="cuda")
U, I = train.shape
eng = get_engine(0)
ptr = torch.from_numpy(train.indptr.astype(np.int64)).cuda(); idx = torch.from_numpy(train.indices.astype(np.int32)).cuda()
d = ptr[1:] - ptr[:-1]
val = torch.repeat_interleave(1.0 / d.double(), d)
torch.cuda.synchronize()
eng.trace(True)
for rep in range(2):
    t0 = time.time()
    fit = eng.fit_topk_real(U, I, ptr, idx, val, K, similarity=sim)
    eng.sync(); t1 = time.time()
    print(f"[{shape} normalize_X {sim} K={K}] rep{rep} fit {1e3*(t1-t0):.1f} ms; rows with < K neighbours {(fit['len'] < K).sum().item()}", flush=True)
    print(eng.trace_report(), flush=True)
