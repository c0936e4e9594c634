"""Host-side breakdown of one end-to-end pass through the drop-in classes at ML-25M shape (fresh matrix objects per
pass, as bench.py's e2e leg does): wall time per call and a cProfile of one pass."""
import sys, time, cProfile, pstats, io, warnings
sys.path.insert(0, ".")
import numpy as np, torch
from scipy.sparse import csr_matrix
from recpack_b200 import ItemKNN, NDCGK, RecallK
from recpack_b200.synth import make_dataset
train, test_out, _ = make_dataset("ml25m", seed=0, split_seed=42, generator="cuda")
def fresh(M):
    out = csr_matrix((M.data, M.indices, M.indptr), shape=M.shape); out.has_canonical_format = True; return out
def run():
    X, Y = fresh(train), fresh(test_out)
    torch.cuda.synchronize()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0=time.perf_counter(); algo = ItemKNN(K=200, predict_topK=20, remove_history=True).fit(X); t1=time.perf_counter()
        pred = algo.predict(X); t2=time.perf_counter()
    m1, m2 = NDCGK(10), RecallK(20)
    m1.calculate(Y, pred); v1 = m1.value; t3=time.perf_counter(); m2.calculate(Y, pred); v2 = m2.value; t4=time.perf_counter()
    return tuple(round(1e3*x, 2) for x in (t1-t0, t2-t1, t3-t2, t4-t3, t4-t0))
run(); run()
for _ in range(3): print("fit / predict / ndcg / recall / total ms:", run())
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
