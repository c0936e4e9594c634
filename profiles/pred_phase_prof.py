import sys, warnings
sys.path.insert(0, ".")
from bench import make_data
from recpack_b200 import ItemKNN
train, test_out = make_data("ml25m")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    algo = ItemKNN(K=200, predict_topK=20, remove_history=True).fit(train)
    for _ in range(2):
        pred = algo.predict(train)
from recpack_b200.engine import get_engine
print("timings", get_engine().last_timings())
