"""Chunk-length sweep for small/medium n (GPU): kernel time per forced L."""
import json, os, sys
import numpy as np
from scipy.stats import unitary_group
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb
for n in [int(x) for x in sys.argv[1:]]:
    A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
    for _ in range(3):
        pqb.GlynnPermanent(A).calculate()
    best = {}
    os.environ.pop("PQB200_FORCE_L", None)
    ts = []
    for _ in range(5):
        pqb.GlynnPermanent(A).calculate(); ts.append(pqb.last_stats()["kernel_ms"])
    best["default"] = min(ts)
    for L in range(2, min(n - 1, 12) + 1):
        os.environ["PQB200_FORCE_L"] = str(L)
        ts = []
        for _ in range(5):
            pqb.GlynnPermanent(A).calculate(); ts.append(pqb.last_stats()["kernel_ms"])
        best[L] = min(ts)
    os.environ.pop("PQB200_FORCE_L", None)
    print(json.dumps({"n": n, "kernel_ms_by_L": {str(k): round(v, 4) for k, v in best.items()}}), flush=True)
