#!/usr/bin/env python
"""Times one fit -> model load -> top-N predict of a named shape on cuda:0 and prints the kernel timings.
usage: probe_config.py <shape> <similarity> <K> [dense_users]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recpack_b200.engine import get_engine
from recpack_b200.synth import make_dataset

shape, sim, K = sys.argv[1], sys.argv[2], int(sys.argv[3])
t0 = time.time()
train, test_out, gen = make_dataset(shape, generator="auto")
U, I = train.shape
d = np.diff(train.indptr); n = np.bincount(train.indices, minlength=I)
print(f"[{shape}] data {time.time()-t0:.1f}s gen={gen} U={U} I={I} nnz={train.nnz} max_d={d.max()} max_n={n.max()} heavy_items={(n>=65536).sum()} sum_d2={(d.astype(float)**2).sum():.3e}", flush=True)
eng = get_engine(0)
if len(sys.argv) > 4:
    eng.fit_config(int(sys.argv[4]))
ptr = torch.from_numpy(train.indptr.astype(np.int64)).cuda(); idx = torch.from_numpy(train.indices.astype(np.int32)).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    fit = eng.fit_topk(U, I, ptr, idx, K, similarity=sim, want_cnt=False)
    eng.sync(); t1 = time.time()
    eng.model_load_topk(I, K, fit["idx"], fit["val"], fit["len"])
    eng.sync(); t2 = time.time()
    top = eng.predict_topn(U, ptr, idx, 20, mask_history=True, want_val=False)
    eng.sync(); t3 = time.time()
    print(f"[{shape}] rep{rep} fit {1e3*(t1-t0):.1f} ms  model {1e3*(t2-t1):.1f} ms  predict {1e3*(t3-t2):.1f} ms  kernels {eng.last_timings()}", flush=True)
print(f"[{shape}] rows with < K neighbours: {(fit['len'] < K).sum().item()}, mean list len {top['len'].float().mean().item():.2f}, mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB torch", flush=True)
