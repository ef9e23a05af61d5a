import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
U = unitary_group.rvs(60, random_state=4)
inp = np.zeros(60, dtype=np.int64); inp[:36] = 1
out = np.zeros(60, dtype=np.int64); out[:18] = 2
for _ in range(3):
    v = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
    print(v, pqb.last_stats())
