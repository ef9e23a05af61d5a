set -x
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
rng = np.random.default_rng(1)
V = np.ascontiguousarray(unitary_group.rvs(14, random_state=5))
st = np.array([1] * 10 + [0] * 4, dtype=np.int64)
for shots in (3, 30, 460):      # clusters of 8 / 8 / two-launch scheme
    s = pqb.GeneralizedCliffordsBSimulationStrategy(V, seed=5).simulate(st, shots)
smp = np.array([rng.multinomial(5, [1 / 14] * 14) for _ in range(20)], dtype=np.int64)
cur = np.array([rng.multinomial(6, [1 / 14] * 14) for _ in range(20)], dtype=np.int64)
pqb.compute_pmf(V, smp, cur); pqb.compute_pmf(V, smp, cur, batched=True)
U = np.ascontiguousarray(unitary_group.rvs(10, random_state=3))
for photons in (6, 12):
    inp = rng.multinomial(photons, rng.dirichlet(np.ones(10))).astype(np.int64)
    out = rng.multinomial(photons, rng.dirichlet(np.ones(10) * 0.3)).astype(np.int64)
    pqb.BBFGRepeatedPermanentCalculatorDouble(U, inp, out).calculate()
pqb.GlynnPermanent(np.ascontiguousarray(unitary_group.rvs(18, random_state=2))).calculate()
ins = [rng.multinomial(5, [1 / 10] * 10).astype(np.int64) for _ in range(10)]
pqb.ChinHuhPermanentCalculator_wrapper(7, U, ins, ins).calculate()
print("racecheck workload done")
PY
compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python /tmp/san2.py > gpurun_out/sanitizer_racecheck_r2.log 2>&1; echo "racecheck rc $?"
tail -8 gpurun_out/sanitizer_racecheck_r2.log
compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python /tmp/san2.py > gpurun_out/sanitizer_synccheck_r2.log 2>&1; echo "synccheck rc $?"
tail -4 gpurun_out/sanitizer_synccheck_r2.log
