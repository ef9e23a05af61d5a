set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
from oracle import oraclebind
# plain, a few sizes (static + rolled forms)
for n in (3, 8, 16, 20, 27, 31):
    A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
    v = pqb.GlynnPermanent(A).calculate()
    if n <= 20:
        w = oraclebind.bbfg(A, oraclebind.LONG_DOUBLE); assert abs(v - w) <= 1e-10 * abs(w), (n, v, w)
# repeated: both kernel forms, several C
rng = np.random.default_rng(1)
U = np.ascontiguousarray(unitary_group.rvs(12, random_state=3))
for photons in (4, 9, 14, 30, 60):
    for t in range(3):
        inp = rng.multinomial(photons, rng.dirichlet(np.ones(12))).astype(np.int64)
        out = rng.multinomial(photons, rng.dirichlet(np.ones(12) * 0.3)).astype(np.int64)
        v = pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
# double-double modes
pqb.permanent_double_double(np.ascontiguousarray(unitary_group.rvs(10, random_state=1)))
pqb.permanent_repeated_double_double(U, [3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0])
# batch forms
ins = [rng.multinomial(5, [1 / 12] * 12).astype(np.int64) for _ in range(20)]
pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, ins).calculate()
pqb.BBFGPermanentDouble([np.ascontiguousarray(unitary_group.rvs(k, random_state=k)) for k in (2, 5, 9, 13)]).calculate()
# sampler: cluster path (few shots), single-CTA path, two-launch path, pmf layer, lossy strategies
V = np.ascontiguousarray(unitary_group.rvs(16, random_state=5))
st = np.array([1] * 12 + [0] * 4, dtype=np.int64)
for shots in (3, 40, 500):
    s = pqb.GeneralizedCliffordsBSimulationStrategy(V, seed=5).simulate(st, shots)
    assert all(x.sum() == 12 for x in s)
pqb.GeneralizedCliffordsBUniformLossesSimulationStrategy(V, 0.8, seed=5).simulate(st, 100)
pqb.GeneralizedCliffordsBLossySimulationStrategy(V * 0.9, 2, seed=5).simulate(st, 100)
smp = np.array([rng.multinomial(5, [1 / 16] * 16) for _ in range(50)], dtype=np.int64)
cur = np.array([rng.multinomial(6, [1 / 16] * 16) for _ in range(50)], dtype=np.int64)
pqb.compute_pmf(V, smp, cur); pqb.compute_pmf(V, smp, cur, batched=True)
print("sanitizer workload done")
PY
compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_memcheck_r2.log 2>&1; echo "memcheck rc $?"
tail -6 gpurun_out/sanitizer_memcheck_r2.log
