#!/usr/bin/env python
"""Emulates every rank's scoring and fit shard of an N-GPU run on ONE GPU (the kernels are the same; only the exchange
is missing) and times them: shows the load balance the sharding rule actually achieves.  usage: probe_shards.py <world>"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recpack_b200.engine import get_engine
from recpack_b200.synth import make_dataset
from recpack_b200.distributed import shard_bounds, fit_work_per_item, score_work_per_user

world = int(sys.argv[1])
train, test_out, _ = make_dataset("ml25m", seed=0, split_seed=42, generator="numpy")
U, I = train.shape
K, N = 200, 20
eng = get_engine(0); eng.use_torch_stream()
dev = torch.device("cuda", 0)
ptr = torch.from_numpy(train.indptr.astype(np.int64)).to(dev); idx = torch.from_numpy(train.indices.astype(np.int32)).to(dev)
fit = eng.fit_topk(U, I, ptr, idx, K, want_cnt=False)
eng.model_load_topk(I, K, fit["idx"], fit["val"], fit["len"])
icut = shard_bounds(fit_work_per_item(train), world)
ucut = shard_bounds(score_work_per_user(train, K), world)
def timed(fn, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
for r in range(world):
    ub, ue = ucut[r], ucut[r + 1]
    p_h = train.indptr[ub:ue + 1].astype(np.int64); lo, hi = int(p_h[0]), int(p_h[-1])
    u_ptr = torch.from_numpy(p_h - lo).to(dev); u_idx = idx[lo:hi].contiguous()
    out = {"idx": torch.empty((ue - ub, N), dtype=torch.int32, device=dev), "val": None, "len": torch.empty((ue - ub,), dtype=torch.int32, device=dev)}
    t_pred = timed(lambda: eng.predict_topn(ue - ub, u_ptr, u_idx, N, mask_history=True, out=out))
    kp = eng.last_timings()["predict_ms"]
    ib, ie = icut[r], icut[r + 1]
    t_fit = timed(lambda: eng.fit_topk(U, I, ptr, idx, K, item_begin=ib, item_end=ie, want_cnt=False))
    kt = eng.last_timings()
    print(f"rank {r}/{world}: users {ue-ub} hist {hi-lo} predict {t_pred:.3f} ms (kernel {kp:.3f})   rows {ie-ib} fit {t_fit:.3f} ms (gram {kt['gram_tc_ms']:.3f} rows {kt['fit_rows_ms']:.3f})", flush=True)
