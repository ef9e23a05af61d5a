"""A few launches of the repeated-row kernel on BASELINE config C4 (30 photons / 60 modes, multinomial
occupation): the workload of the ncu capture in profiles/."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb

U = unitary_group.rvs(60, random_state=4)
inp = np.zeros(60, dtype=np.int64); inp[:30] = 1
out = np.random.default_rng(4).multinomial(30, [1 / 60] * 60).astype(np.int64)
for r in range(3):
    t = time.time()
    v = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
    st = pqb.last_stats()
    print(f"C4 multinomial: {v} kernel {st['kernel_ms']:.3f} ms, {st['gray_steps']:.4g} steps, "
          f"{st['gray_steps'] / st['kernel_ms'] * 1e3:.4e} steps/s, ndev {st['ndev']}, wall {1e3 * (time.time() - t):.2f} ms")
