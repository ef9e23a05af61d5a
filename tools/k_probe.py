"""Sub-chunk length sweep (GPU): kernel time and relative error vs the committed double-double value.
usage: python tools/k_probe.py n K [K ...]   (K == chunk length 12 means: no anchor re-seeding)"""
import json, os, sys
import numpy as np
from scipy.stats import unitary_group
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "plain_dd_reference.json")))["cases"][str(n)]
g = complex(*[float.fromhex(x) for x in gold["dd_parts_hex"]][0::2])
A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
pqb.GlynnPermanent(A[:n - 8, :n - 8]).calculate()
for K in sys.argv[2:]:
    os.environ["PQB200_FORCE_K"] = K
    v = pqb.GlynnPermanent(A).calculate()
    st = pqb.last_stats()
    print(json.dumps({"n": n, "K": int(K), "kernel_ms": st["kernel_ms"], "rel_err_vs_dd": abs(v - g) / abs(g)}), flush=True)
