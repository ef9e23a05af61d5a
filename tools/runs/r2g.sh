set -x
mkdir -p gpurun_out
python -m pytest tests/test_reference_wrappers_gpu.py -m gpu -q --timeout=900 2>&1 | tail -15
# sampler A/B against the round-1 library and package (git worktree r1ref of 74fc7eb)
cat > /tmp/ab.py <<'PY'
import os, sys, time, json
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
print(pqb.__file__)
U = np.ascontiguousarray(unitary_group.rvs(40, random_state=5))
state = np.array([1] * 20 + [0] * 20, dtype=np.int64)
for shots in (125, 1000, 3000, 20000):
    sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=5, ndev=1)
    sim.simulate(state, shots)
    best = (1e9, None)
    for r in range(3):
        sim.seed(5)
        t = time.perf_counter(); res = sim.simulate(state, shots); dt = time.perf_counter() - t
        st = pqb.last_stats()
        if dt < best[0]: best = (dt, st)
    print(json.dumps({"shots": shots, "call_ms": best[0] * 1e3, "kernel_ms": best[1]["kernel_ms"], "launches": best[1]["launches"]}))
PY
echo "== round 1 build"; PYTHONPATH=$PWD/r1ref python /tmp/ab.py
echo "== current build"; PYTHONPATH=$PWD python /tmp/ab.py
echo "== current build, chunks 4"; PQB200_LAP_CHUNKS=4 PYTHONPATH=$PWD python /tmp/ab.py
