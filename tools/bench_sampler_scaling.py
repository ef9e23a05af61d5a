"""C5-shaped sampling (20 photons / 40 modes) with many shots, sharded over 1..ndev GPUs of one process
(shots shard with no collective: every device finishes whole shots)."""
import json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
from piquassoboost_b200 import _lib

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
U = unitary_group.rvs(40, random_state=5)
inp = np.zeros(40, dtype=np.int64)
inp[random.Random(5).sample(range(40), 20)] = 1
ndev_max = _lib.load().pqb200_device_count()
ref = None
n = 1
while n <= ndev_max:
    s = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=123, ndev=n)
    s.simulate(inp, 200)
    s.seed(123)
    t0 = time.perf_counter()
    got = np.array(s.simulate(inp, shots))
    t = time.perf_counter() - t0
    st = pqb.last_stats()
    if ref is None:
        ref = got
    print(json.dumps({"ndev": n, "shots": shots, "wall_s": t, "us_per_shot": 1e6 * t / shots, "kernel_ms_dev0": st["kernel_ms"],
                      "identical_to_1_gpu": bool((got == ref).all())}), flush=True)
    n *= 2
