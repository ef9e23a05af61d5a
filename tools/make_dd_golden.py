"""Assemble tests/golden/plain_dd_reference.json.

Inputs (both produced by committed scripts):
  * the JSON lines of `python tools/accuracy_probe.py 40 30` run on a B200 (gpurun): per n the
    double-double evaluation pqb200_permanent_dd of scipy.stats.unitary_group.rvs(n, random_state=n)
    as four hex doubles, and for n <= 30 the unmodified reference's long double value computed in the
    same run (oracle/_ref);
  * the JSON written by `python tools/ref_big.py 32 34 36` in the dev container: the unmodified
    reference's long double BBFG calculator on the same matrices (96 s / 423 s / 1861 s on 8 cores; `tools/ref_big.py 38`: 7594 s).

usage: python tools/make_dd_golden.py gpurun_out/accuracy.jsonl tests/golden/inputs/ref_big_32_34_36.json tests/golden/inputs/ref_big_38.json
(the "40_sub80" case is appended from the output of tools/accuracy_subblock.py 40)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(acc_path, *ref_paths):
    out = {"_comment": "Double-double (106-bit) GPU evaluation pqb200_permanent_dd of "
                       "scipy.stats.unitary_group.rvs(n, random_state=n); dd_parts_hex = re.hi re.lo im.hi im.lo. "
                       "ref_long_double = the unmodified reference's long double BBFG calculator on the same matrix "
                       "(n <= 30: same GPU-box run; n = 32..36: tools/ref_big.py on CPU). Made by tools/make_dd_golden.py.",
           "cases": {}}
    for line in open(acc_path):
        r = json.loads(line)
        if "dd" not in r:
            continue
        c = {"matrix": r["matrix"], "dd_parts_hex": r["dd"], "dd_value": r["dd_value"]}
        if "ref_long_double" in r:
            c["ref_long_double"] = r["ref_long_double"]
            c["dd_rel_err_vs_ref"] = r["dd_rel_err_vs_ref"]
        out["cases"][str(r["n"])] = c
    refs = {}
    for ref_path in ref_paths:
        refs.update(json.load(open(ref_path)))
    for n, r in refs.items():
        if n not in out["cases"]:
            continue
        v = complex(r["re_f"], r["im_f"])
        d = complex(*out["cases"][n]["dd_value"])
        out["cases"][n]["ref_long_double"] = [r["re_f"], r["im_f"]]
        out["cases"][n]["ref_long_double_seconds_8_cores"] = r["seconds"]
        out["cases"][n]["dd_rel_err_vs_ref"] = abs(d - v) / abs(v)
    path = os.path.join(ROOT, "tests", "golden", "plain_dd_reference.json")
    if os.path.exists(path):  # keep cases made by other tools (e.g. "40_sub80" from tools/accuracy_subblock.py)
        for key, case in json.load(open(path))["cases"].items():
            out["cases"].setdefault(key, case)
    json.dump(out, open(path, "w"), indent=1)
    print({n: c.get("dd_rel_err_vs_ref") for n, c in out["cases"].items()})


if __name__ == "__main__":
    main(sys.argv[1], *sys.argv[2:])
