cat > /tmp/ab.py <<'PY'
import os, sys, time, json
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
U = np.ascontiguousarray(unitary_group.rvs(40, random_state=5))
state = np.array([1] * 20 + [0] * 20, dtype=np.int64)
out = []
for shots in (1000, 20000):
    sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=5, ndev=1)
    sim.simulate(state, shots)
    best = (1e9, None)
    for r in range(5):
        sim.seed(5)
        t = time.perf_counter(); res = sim.simulate(state, shots); dt = time.perf_counter() - t
        st = pqb.last_stats()
        if st["kernel_ms"] < best[0]: best = (st["kernel_ms"], dt)
    out.append("%d shots: kernel %.3f ms call %.3f ms" % (shots, best[0], best[1] * 1e3))
print(os.environ.get("PQB200_LIB", "default")[-28:], " | ".join(out))
PY
export PQB200_LAP_TAIL=0 PQB200_LAP_WAVES=1
PYTHONPATH=$PWD/r1ref python /tmp/ab.py
for rep in 1 2; do
for v in "" _SCAN _NOCLUSTER _SCANNOCLUSTER; do PQB200_LIB=$PWD/piquassoboost_b200/libpqb200$v.so PYTHONPATH=$PWD python /tmp/ab.py; done
done
PYTHONPATH=$PWD/r1ref python /tmp/ab.py
