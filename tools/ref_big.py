"""Run the unmodified reference's long-double BBFG calculator on big Haar unitaries (CPU, slow) and
record the values as golden vectors."""
import json, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from scipy.stats import unitary_group
from oracle import refbind
out = {}
for n in [int(x) for x in sys.argv[1:]]:
    A = np.ascontiguousarray(unitary_group.rvs(n, random_state=n))
    t = time.time()
    v = refbind.bbfg(A, True)
    dt = time.time() - t
    out[str(n)] = {"re": v.real.hex(), "im": v.imag.hex(), "re_f": v.real, "im_f": v.imag, "seconds": dt,
                   "matrix": f"scipy.stats.unitary_group.rvs({n}, random_state={n})"}
    print(n, v, dt, flush=True)
    json.dump(out, open(f"/root/repo/scratch/ref_big_{'_'.join(sys.argv[1:])}.json", "w"), indent=1)
