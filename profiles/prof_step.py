import sys, warnings
sys.path.insert(0, ".")
from bench import make_data
from recpack_b200 import ItemKNN, NDCGK, RecallK
train, test_out = make_data("ml25m")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(2):
        algo = ItemKNN(K=200, predict_topK=20, remove_history=True).fit(train)
        pred = algo.predict(train)
        NDCGK(10).calculate(test_out, pred); RecallK(20).calculate(test_out, pred)
