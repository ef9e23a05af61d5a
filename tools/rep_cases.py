"""Kernel times of a few repeated-row permanents of different N (tuning aid; PQB200_LIB selects the build)."""
import os, sys
os.environ.setdefault("PQB200_NO_DD_FALLBACK", "1")   # time the fast kernel only
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb

def run(name, U, inp, out):
    best = 1e9
    for _ in range(4):
        v = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
        st = pqb.last_stats(); best = min(best, st["kernel_ms"])
    print(f"{os.environ.get('PQB200_LIB', 'default')[-12:]} {name}: N={int(np.sum(inp))} steps={st['gray_steps']:.3g} kernel {best:.4f} ms {st['gray_steps']/best*1e3:.3e} steps/s")

U = unitary_group.rvs(60, random_state=4)
inp = np.zeros(60, dtype=np.int64); inp[:30] = 1
run("C4-multinomial", U, inp, np.random.default_rng(4).multinomial(30, [1 / 60] * 60).astype(np.int64))
out = np.zeros(60, dtype=np.int64); out[:15] = 2
run("C4-15x2", U, inp, out)
inp = np.zeros(60, dtype=np.int64); inp[:40] = 1
out = np.zeros(60, dtype=np.int64); out[:12] = 3; out[12:14] = 2
run("N40-12x3+2x2", U, inp, out)
inp = np.zeros(60, dtype=np.int64); inp[:20] = 1
out = np.zeros(60, dtype=np.int64); out[:10] = 2
run("N20-10x2", U, inp, out)
inp = np.zeros(60, dtype=np.int64); inp[:36] = 1
out = np.zeros(60, dtype=np.int64); out[:18] = 2
run("N36-18x2", U, inp, out)
# both sides repeated: the power form keeps C < N column sums
inp = np.zeros(60, dtype=np.int64); inp[:10] = 3
out = np.zeros(60, dtype=np.int64); out[:15] = 2
run("N30-10x3-into-15x2", U, inp, out)
inp = np.zeros(60, dtype=np.int64); inp[:15] = 2
out = np.zeros(60, dtype=np.int64); out[20:35] = 2
run("N30-15x2-into-15x2", U, inp, out)
inp = np.zeros(60, dtype=np.int64); inp[:10] = 6
out = np.zeros(60, dtype=np.int64); out[:10] = [7, 5, 6, 6, 8, 4, 6, 6, 5, 7]
run("N60-10x6-into-10", U, inp, out)
