"""TEST INFRASTRUCTURE -- ``import ot`` must succeed (``src/utils/ot_utils.py:1``).  POT 0.7.0's
``emd`` (exact network simplex) is a training-only loss outside the forward hot path; the stand-in
solves the same transport LP with scipy's HiGHS (the optimal value is unique)."""
import numpy as np


def emd(a, b, M, numItermax=100000):
    from scipy.optimize import linprog
    n, m = M.shape
    A_eq = np.zeros((n + m, n * m))
    for i in range(n):
        A_eq[i, i * m:(i + 1) * m] = 1.0
    for j in range(m):
        A_eq[n + j, j::m] = 1.0
    res = linprog(M.reshape(-1), A_eq=A_eq[:-1], b_eq=np.concatenate([a, b])[:-1], bounds=(0, None), method='highs')
    return res.x.reshape(n, m)
