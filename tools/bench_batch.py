"""Many tiny repeated permanents in one call (the connector / strategy-A regime, SURVEY.md section 8(f)
rank 3-4): every 6-photon output state of an 8-mode interferometer against one input state, through
the list branch of ChinHuhPermanentCalculator_wrapper (one pqb200_permanent_repeated_batch call), next
to a loop over the unmodified reference's BBFG repeated calculator on the host."""
import itertools, json, os, sys, time
import numpy as np
from scipy.stats import unitary_group
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb
from oracle import refbind

m, photons = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 6)
U = unitary_group.rvs(m, random_state=m)
outs = []
for comb in itertools.combinations_with_replacement(range(m), photons):
    s = np.zeros(m, dtype=np.int64)
    for k in comb:
        s[k] += 1
    outs.append(s)
inp = np.zeros(m, dtype=np.int64); inp[:photons] = 1
calc = pqb.ChinHuhPermanentCalculator_wrapper(7, U, [inp], outs)
calc.calculate()
t0 = time.perf_counter(); res = calc.calculate()[0]; t = time.perf_counter() - t0
st = pqb.last_stats()
rec = {"modes": m, "photons": photons, "permanents": len(outs), "gpu_call_s": t, "gpu_us_per_permanent": 1e6 * t / len(outs),
       "kernel_ms": st["kernel_ms"], "launches": st["launches"]}
if refbind.available():
    n = min(len(outs), 400)
    t0 = time.perf_counter()
    want = [refbind.bbfg_repeated(U, inp, o, False) for o in outs[:n]]
    tr = time.perf_counter() - t0
    rec["ref_double_us_per_permanent"] = 1e6 * tr / n
    rec["max_rel_err_vs_ref_double"] = max(abs(a - b) / abs(b) for a, b in zip(res[:n], want) if abs(b) > 1e-14)
    rec["prob_sum"] = None
print(json.dumps(rec))
