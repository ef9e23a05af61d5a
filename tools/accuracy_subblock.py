"""n=40 Haar SUB-matrix (top-left 40x40 of unitary_group.rvs(80, random_state=80), SURVEY.md section 8(d)
config C3): the fast kernel against the double-double evaluation, both on the GPU."""
import json, os, sys, time
import numpy as np
from scipy.stats import unitary_group
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb
from piquassoboost_b200.Boson_Sampling_Utilities import permanent_double_double

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
A = np.ascontiguousarray(unitary_group.rvs(2 * n, random_state=2 * n)[:n, :n])
pqb.GlynnPermanent(A[:20, :20]).calculate()
t = time.time(); fast = pqb.GlynnPermanent(A).calculate(); tf = time.time() - t
t = time.time(); ddv, parts = permanent_double_double(A); td = time.time() - t
print(json.dumps({"n": n, "matrix": f"unitary_group.rvs({2*n}, random_state={2*n})[:{n}, :{n}]", "fast": [fast.real, fast.imag],
                  "fast_s": tf, "dd_s": td, "dd_parts_hex": [float(x).hex() for x in parts], "dd_value": [ddv.real, ddv.imag],
                  "fast_rel_err_vs_dd": abs(fast - ddv) / abs(ddv)}))
