"""Add a reference-held value to tests/golden/plain_dd_reference.json in place.

usage: python tools/add_ref_big.py scratch/ref_big_40.json
The input is what `python tools/ref_big.py <n>` writes in the dev container (the UNMODIFIED reference's long double
BBFG calculator on scipy.stats.unitary_group.rvs(n, random_state=n)); it is copied to tests/golden/inputs/ and the
case n of the golden file gets ref_long_double, the run time and the distance of the committed double-double value
from it.  (tools/make_dd_golden.py rebuilds the whole file from the GPU run's output; this only appends.)"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path):
    refs = json.load(open(path))
    gpath = os.path.join(ROOT, "tests", "golden", "plain_dd_reference.json")
    g = json.load(open(gpath))
    for n, r in refs.items():
        case = g["cases"][n]
        v = complex(r["re_f"], r["im_f"])
        assert float.fromhex(r["re"]) == r["re_f"] and float.fromhex(r["im"]) == r["im_f"]
        d = complex(*case["dd_value"])
        case["ref_long_double"] = [r["re_f"], r["im_f"]]
        case["ref_long_double_seconds_8_cores"] = r["seconds"]
        case["dd_rel_err_vs_ref"] = abs(d - v) / abs(v)
        print(n, "dd vs reference long double:", case["dd_rel_err_vs_ref"], "reference run:", r["seconds"], "s")
    json.dump(g, open(gpath, "w"), indent=1)
    dst = os.path.join(ROOT, "tests", "golden", "inputs", os.path.basename(path))
    if os.path.abspath(path) != dst:
        shutil.copy(path, dst)


if __name__ == "__main__":
    main(sys.argv[1])
