"""Kernel / call time of the C5-shaped B sampler (20 photons, 40 modes) for a few shot counts under the
tuning knobs PQB200_LAP_CLUSTER / PQB200_LAP_WAVES / PQB200_LAP_CHUNKS (one child process per setting: the
knobs are read from the environment)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, time, json
sys.path.insert(0, %r)
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
U = np.ascontiguousarray(unitary_group.rvs(40, random_state=5))
state = np.array([1] * 20 + [0] * 20, dtype=np.int64)
shots = int(sys.argv[1]); nd = int(sys.argv[2])
sim = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=5, ndev=nd)
sim.simulate(state, shots)
best = (1e9, None)
for r in range(int(os.environ.get('SWEEP_REPS', '3'))):
    sim.seed(5)
    t = time.perf_counter(); res = sim.simulate(state, shots); dt = time.perf_counter() - t
    st = pqb.last_stats()
    if dt < best[0]: best = (dt, st)
print(json.dumps({"shots": shots, "ndev": best[1]["ndev"], "call_ms": best[0] * 1e3, "kernel_ms": best[1]["kernel_ms"],
                  "launches": best[1]["launches"], "checksum": int(np.array(res).astype(np.int64).dot(np.arange(40)).sum())}))
""" % ROOT
ndevs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
shot_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "125,1000,20000").split(",")]
if len(sys.argv) > 3:                             # explicit list of knob settings (JSON)
    settings = json.loads(sys.argv[3])
else:
    settings = [{}]                               # the library's own choice
    for tail in ("0", "2", "3", "4", "5"):
        settings.append({"PQB200_LAP_TAIL": tail})
    settings.append({"PQB200_LAP_TAIL": "3", "PQB200_LAP_WAVES": "1"})
for shots in shot_list:
    for nd in ndevs:
        for st in settings:
            env = dict(os.environ)
            env.update(st)
            r = subprocess.run([sys.executable, "-c", CHILD, str(shots), str(nd)], capture_output=True, text=True, env=env)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(f"{st or 'default'} {line}", flush=True)
