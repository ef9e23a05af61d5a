"""Full-size known-answer run at the largest supported n (in-process multi-GPU, `ndev` = all devices):
perm(A (+) B) = perm(A) perm(B) for two Haar blocks, the factors computed by the unmodified reference in
long double.  usage: python tools/big_check.py 45"""
import json, os, sys, time
import numpy as np
from scipy.stats import unitary_group
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb
from oracle import refbind

n = int(sys.argv[1]) if len(sys.argv) > 1 else 45
a = n // 2
A, B = unitary_group.rvs(a, random_state=1000 + a), unitary_group.rvs(n - a, random_state=2000 + n - a)
M = np.zeros((n, n), dtype=complex)
M[:a, :a] = A
M[a:, a:] = B
pqb.GlynnPermanent(M[:20, :20]).calculate()
t0 = time.perf_counter()
got = pqb.GlynnPermanent(M).calculate()
t = time.perf_counter() - t0
st = pqb.last_stats()
want = refbind.bbfg(A, True) * refbind.bbfg(B, True)
print(json.dumps({"n": n, "blocks": [a, n - a], "ndev": st["ndev"], "wall_s": t, "kernel_ms": st["kernel_ms"],
                  "gray_steps": st["gray_steps"], "steps_per_s": st["gray_steps"] / t,
                  "rel_err_vs_reference_long_double_product": abs(got - want) / abs(want),
                  "value": [got.real, got.imag]}))
