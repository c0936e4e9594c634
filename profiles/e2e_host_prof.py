import sys, time, cProfile, pstats, io, warnings
sys.path.insert(0, ".")
import numpy as np
from bench import make_data
from recpack_b200 import ItemKNN, NDCGK, RecallK
train, test_out = make_data("ml25m")
def run():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0=time.perf_counter(); algo = ItemKNN(K=200, predict_topK=20, remove_history=True).fit(train); t1=time.perf_counter()
        pred = algo.predict(train); t2=time.perf_counter()
    m1, m2 = NDCGK(10), RecallK(20)
    m1.calculate(test_out, pred); t3=time.perf_counter(); m2.calculate(test_out, pred); t4=time.perf_counter()
    return t1-t0, t2-t1, t3-t2, t4-t3
run(); print("fit/predict/ndcg/recall s:", run())
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
