"""Two fit -> predict -> metrics passes at ML-25M shape through the drop-in classes; run with
RPK_LIB=profiles/librpk_prof.so (a -DRPK_PHASE_PROF build) to get the per-phase cycle tables on stderr."""
import sys, warnings
sys.path.insert(0, ".")
from recpack_b200 import ItemKNN, NDCGK, RecallK
from recpack_b200.synth import make_dataset
train, test_out, _ = make_dataset("ml25m", seed=0, split_seed=42, generator="cuda")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(2):
        algo = ItemKNN(K=200, predict_topK=20, remove_history=True).fit(train)
        pred = algo.predict(train)
        NDCGK(10).calculate(test_out, pred); RecallK(20).calculate(test_out, pred)
