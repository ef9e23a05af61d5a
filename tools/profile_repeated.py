"""A few launches of the repeated-row kernel for the ncu captures in profiles/:
  c4  : BASELINE config C4 (30 photons / 60 modes, multinomial output, single inputs) -> glynn_reppow_kernel<30,false>
  n60 : 60 photons in 10 modes on both sides (power form)                             -> glynn_reppow_kernel<10,true>"""
import os, sys, time
os.environ.setdefault("PQB200_NO_DD_FALLBACK", "1")   # profile the fast kernel only
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb

case = sys.argv[1] if len(sys.argv) > 1 else "c4"
U = unitary_group.rvs(60, random_state=4)
if case == "c4":
    inp = np.zeros(60, dtype=np.int64); inp[:30] = 1
    out = np.random.default_rng(4).multinomial(30, [1 / 60] * 60).astype(np.int64)
else:
    inp = np.zeros(60, dtype=np.int64); inp[:10] = 6
    out = np.zeros(60, dtype=np.int64); out[:10] = [7, 5, 6, 6, 8, 4, 6, 6, 5, 7]
for r in range(3):
    t = time.time()
    v = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
    st = pqb.last_stats()
    print(f"{case}: {v} kernel {st['kernel_ms']:.3f} ms, {st['gray_steps']:.4g} steps, "
          f"{st['gray_steps'] / st['kernel_ms'] * 1e3:.4e} steps/s, ndev {st['ndev']}, wall {1e3 * (time.time() - t):.2f} ms")
