"""Run a few launches of the n=40 plain kernel on 1/512 of the Gray space (2^30 steps each): the
workload used for the ncu captures in profiles/ (a full n=40 permanent is ~8 s per launch, far too
long under ncu's ~40 replays)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
from piquassoboost_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
lib = _lib.load()
out = np.zeros(4)
for r in range(reps):
    t = time.time()
    _lib.check(lib.pqb200_permanent_shard(_lib.dptr(A), n, n, n, 0, shards // 2, shards, _lib.dptr(out), None, None))
    st = _lib.last_stats()
    print(f"n={n} shard 1/{shards}: kernel {st['kernel_ms']:.3f} ms, {st['gray_steps']/st['kernel_ms']*1e3:.4e} steps/s, "
          f"{st['gray_steps']*8*n/st['kernel_ms']/1e9:.2f} TFLOP/s, wall {1e3*(time.time()-t):.1f} ms", out)
