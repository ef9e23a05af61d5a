"""Accuracy probe (GPU): the fast kernel against the double-double evaluation and, where it is
affordable, the unmodified reference's long double calculator.  Prints one JSON line per case.
usage: python tools/accuracy_probe.py [max_n_dd=40] [ref_max_n=30]"""
import json
import os
import sys
import time

import numpy as np
from scipy.stats import unitary_group

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb  # noqa: E402
from piquassoboost_b200.Boson_Sampling_Utilities import permanent_double_double  # noqa: E402
from oracle import refbind  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                   "plain_dd_reference.json")))["cases"]
max_dd = int(sys.argv[1]) if len(sys.argv) > 1 else 40  # largest n the dd kernel is RE-RUN for
ref_max = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pqb.load()


def rel(a, b):
    return abs(a - b) / abs(b)


dd_rate = None
for n in [20, 24, 28, 30, 32, 34, 36, 38, 40]:
    A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
    rec = {"n": n, "matrix": f"unitary_group.rvs({n}, random_state={n})"}
    t = time.time()
    fast = pqb.GlynnPermanent(A).calculate()
    rec["fast_s"] = time.time() - t
    rec["fast"] = [fast.real, fast.imag]
    for L in (6, 8, 10):
        if L < n - 1:
            os.environ["PQB200_FORCE_L"] = str(L)
            v = pqb.GlynnPermanent(A).calculate()
            rec[f"fast_L{L}_rel_vs_default"] = rel(v, fast)
            rec[f"fast_L{L}"] = [v.real, v.imag]
    os.environ.pop("PQB200_FORCE_L", None)
    if str(n) in GOLD:  # committed double-double value (and reference long double where it exists)
        gold = complex(*[float.fromhex(x) for x in GOLD[str(n)]["dd_parts_hex"]][0::2])
        rec["fast_rel_err_vs_golden_dd"] = rel(fast, gold)
        for L in (6, 8, 10):
            if f"fast_L{L}" in rec:
                rec[f"fast_L{L}_rel_err_vs_golden_dd"] = rel(complex(*rec[f"fast_L{L}"]), gold)
    est = None if dd_rate is None else (2.0 ** (n - 1)) * n / dd_rate
    if n <= max_dd and (est is None or est < 420):
        t = time.time()
        ddv, parts = permanent_double_double(A)
        dt = time.time() - t
        dd_rate = (2.0 ** (n - 1)) * n / dt
        rec["dd_s"] = dt
        rec["dd"] = [float(parts[0]).hex(), float(parts[1]).hex(), float(parts[2]).hex(), float(parts[3]).hex()]
        rec["dd_value"] = [ddv.real, ddv.imag]
        rec["fast_rel_err_vs_dd"] = rel(fast, ddv)
        for L in (6, 8, 10):
            if f"fast_L{L}" in rec:
                rec[f"fast_L{L}_rel_err_vs_dd"] = rel(complex(*rec[f"fast_L{L}"]), ddv)
    if n <= ref_max and refbind.available():
        t = time.time()
        ref = refbind.bbfg(A, True)
        rec["ref_long_double_s"] = time.time() - t
        rec["ref_long_double"] = [ref.real, ref.imag]
        rec["fast_rel_err_vs_ref"] = rel(fast, ref)
        if "dd_value" in rec:
            rec["dd_rel_err_vs_ref"] = rel(complex(*rec["dd_value"]), ref)
    print(json.dumps(rec), flush=True)
