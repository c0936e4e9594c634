import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.sparse import csr_matrix
from recpack_b200.engine import get_engine
from recpack_b200.matrix import binary_structure
eng = get_engine(0)
def tryload(name, I, K, idx, val, ln):
    try:
        eng.model_load_topk(I, K, np.ascontiguousarray(idx, dtype=np.int32), np.ascontiguousarray(val, dtype=np.float64), np.ascontiguousarray(ln, dtype=np.int32))
        print(name, "ok")
    except Exception as e:
        print(name, "FAILED", e)
X = csr_matrix(np.array([[1, 1, 0], [1, 1, 0], [0, 0, 1], [0, 0, 1], [0, 0, 0]], dtype=np.int32))
_, indptr, indices = binary_structure(X)
tryload("before fit", 3, 2, [[1,-1],[0,-1],[-1,-1]], [[1.,0.],[1.,0.],[0.,0.]], [1,1,0])
fit = eng.fit_topk(5, 3, indptr, indices, 2)
print([hex(x) for x in fit["val"].view(np.uint64).ravel()], fit["idx"].ravel(), fit["len"])
tryload("after fit, same literal", 3, 2, [[1,-1],[0,-1],[-1,-1]], [[1.,0.],[1.,0.],[0.,0.]], [1,1,0])
tryload("after fit, fit arrays", 3, 2, fit["idx"], fit["val"], fit["len"])
tryload("again", 3, 2, fit["idx"], fit["val"], fit["len"])
v2 = fit["val"].copy(); v2[v2 > 0] = 1.0
tryload("exact 1.0", 3, 2, fit["idx"], v2, fit["len"])
