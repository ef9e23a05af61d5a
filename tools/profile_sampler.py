"""Small C5-shaped sampler run used for the ncu launch list / captures in profiles/."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 200
U = unitary_group.rvs(40, random_state=5)
inp = np.zeros(40, dtype=np.int64)
inp[random.Random(5).sample(range(40), 20)] = 1
pqb.GlynnPermanent(np.eye(4)).calculate()
t = time.perf_counter()
got = pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=123).simulate(inp, shots)
print(f"{shots} shots in {time.perf_counter() - t:.3f} s; stats {pqb.last_stats()}")
