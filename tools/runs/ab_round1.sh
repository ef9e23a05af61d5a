# A/B against the round-1 library ON THE SAME GPU BOX (profiles/rep_cases_vs_round1_r2.log,
# profiles/sampler_vs_round1_r2.log).  In the dev container first:
#     git worktree add -f r1ref 74fc7eb && make -C r1ref/piquassoboost_b200/csrc -j8
# (r1ref/ travels to the GPU box with the snapshot; remove the worktree afterwards), then under gpurun:
echo "== round 1: repeated kernel";  PYTHONPATH=$PWD/r1ref python r1ref/tools/rep_cases.py 2>&1 | head -5
echo "== round 2: repeated kernel";  python tools/rep_cases.py
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
    for r in range(4):
        sim.seed(5)
        t = time.perf_counter(); res = sim.simulate(state, shots); dt = time.perf_counter() - t
        st = pqb.last_stats()
        if dt < best[0]: best = (dt, st)
    print(json.dumps({"shots": shots, "call_ms": best[0] * 1e3, "kernel_ms": best[1]["kernel_ms"], "launches": best[1]["launches"]}))
PY
echo "== round 1: sampler"; PYTHONPATH=$PWD/r1ref python /tmp/ab.py
echo "== round 2: sampler"; PYTHONPATH=$PWD python /tmp/ab.py
