"""Kernel time of one n x n Haar-unitary permanent for every chunk length 2^L the planner may pick
(PQB200_FORCE_L) -- checks the cost model of plan_plain (csrc/pqb200_capi.cu) against measurement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb

for n in [int(x) for x in (sys.argv[1:] or ["30"])]:
    A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
    w = pqb.GlynnPermanentCalculator_wrapper(A, pqb.GlynnB200, ndev=1)
    os.environ.pop("PQB200_FORCE_L", None)
    w.calculate()
    best = min((w.calculate(), pqb.last_stats()["kernel_ms"])[1] for _ in range(5))
    print(f"n={n} planner's choice: {best:.4f} ms  {2.0 ** (n - 1) / best * 1e3:.4e} steps/s")
    for L in range(6, 13):
        if L > n - 1:
            continue
        os.environ["PQB200_FORCE_L"] = str(L)
        best = min((w.calculate(), pqb.last_stats()["kernel_ms"])[1] for _ in range(5))
        print(f"n={n} L={L}: {best:.4f} ms  {2.0 ** (n - 1) / best * 1e3:.4e} steps/s")
    os.environ.pop("PQB200_FORCE_L", None)
