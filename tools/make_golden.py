"""tools/make_golden.py -- regenerate tests/golden/*.json in the DEV CONTAINER (needs /root/reference).

Two sources, both from the reference itself:
  1. its own known-answer tests: unittests/test_permanent.py (6 KATs for
     BBFGRepeatedPermanentCalculatorDouble) and unittests/test_seeding.py (_U matrix + the 20
     lossless seeded shots); extracted by executing those test modules with the piquassoboost /
     piquasso imports stubbed out, so no number is re-typed by hand;
  2. outputs of the UNMODIFIED reference C++ (oracle/_ref/libpqbref.so, built by oracle/Makefile)
     on seeded inputs: long-double BBFG on Haar unitaries / sub-blocks, repeated-row cases with the
     occupation patterns of ctests/permanents_speed.cpp:208-231, compute_pmf rows and the sampler's
     C++ generation order for seed 123.
The GPU box has no /root/reference: tests read only the committed JSON.
"""
import json
import os
import sys
import types

import numpy as np
from scipy.stats import unitary_group

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def c2l(a):
    a = np.asarray(a, dtype=np.complex128)
    return {"shape": list(a.shape), "re": a.real.ravel().tolist(), "im": a.imag.ravel().tolist()}


def z2l(z):
    z = complex(z)
    return [z.real, z.imag]


def kats_from_reference_tests():
    """Run unittests/test_permanent.py with a recording stub in place of the calculator class."""
    src = open(os.path.join(REF, "unittests", "test_permanent.py")).read()
    records = []

    class Recorder:
        def __init__(self, matrix, cols, rows):
            self.args = (np.array(matrix, dtype=np.complex128), np.array(cols), np.array(rows))

        def calculate(self):
            records.append({"args": self.args})
            return ("PENDING", len(records) - 1)

    fake_np = types.ModuleType("numpy_proxy")
    fake_np.__dict__.update(np.__dict__)

    def isclose(value, expected, *a, **k):
        if isinstance(value, tuple) and value[0] == "PENDING":
            records[value[1]]["expected"] = complex(expected)
            return True
        return np.isclose(value, expected, *a, **k)

    fake_np.isclose = isclose
    mod_bsu = types.ModuleType("piquassoboost.sampling.Boson_Sampling_Utilities")
    mod_bsu.BBFGRepeatedPermanentCalculatorDouble = Recorder
    saved = {k: sys.modules.get(k) for k in ("piquassoboost", "piquassoboost.sampling",
                                              "piquassoboost.sampling.Boson_Sampling_Utilities")}
    sys.modules["piquassoboost"] = types.ModuleType("piquassoboost")
    sys.modules["piquassoboost.sampling"] = types.ModuleType("piquassoboost.sampling")
    sys.modules["piquassoboost.sampling.Boson_Sampling_Utilities"] = mod_bsu
    try:
        sys.modules["numpy_proxy"] = fake_np
        g = {"__name__": "ref_test_permanent"}
        exec(compile(src.replace("import numpy as np", "import numpy_proxy as np"), "test_permanent.py", "exec"), g)
        out = []
        for name, fn in g.items():
            if name.startswith("test_") and callable(fn):
                before = len(records)
                fn()
                for rec in records[before:]:
                    m, cols, rows = rec["args"]
                    out.append({"name": name, "matrix": c2l(m), "cols": cols.tolist(),
                                "rows": rows.tolist(), "expected": z2l(rec["expected"]),
                                "source": "unittests/test_permanent.py"})
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop("numpy_proxy", None)


def lossy_expected_from_reference_tests():
    """expected_samples of the two LossyInterferometer tests of unittests/test_seeding.py."""
    import ast
    tree = ast.parse(open(os.path.join(REF, "unittests", "test_seeding.py")).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in (
                "test_LossyInterferometer_boson_sampling_seeded",
                "test_LossyInterferometer_boson_sampling_uniform_losses"):
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and getattr(sub.targets[0], "id", None) == "expected_samples":
                    out[node.name] = [list(t) for t in ast.literal_eval(sub.value)]
    assert len(out) == 2
    return out


def expanded_space(interferometer_svd):
    """piquasso-module/piquasso/_simulators/sampling/utils.py:443-474 (the 2m x 2m unitary dilation
    piquassoboost/sampling/calculations.py:125-135 feeds to the lossless B strategy), executed from
    the reference source itself."""
    import ast
    src = open(os.path.join(REF, "piquasso-module", "piquasso", "_simulators", "sampling", "utils.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in (
        "_prepare_interferometer_matrix_in_expanded_space", "_calculate_singular_values_matrix_expansion")]
    g = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "utils.py", "exec"), g)
    return g["_prepare_interferometer_matrix_in_expanded_space"](interferometer_svd)


def seeding_from_reference_tests():
    """_U and the expected lossless shots of unittests/test_seeding.py (seed 123)."""
    src = open(os.path.join(REF, "unittests", "test_seeding.py")).read()
    import ast
    tree = ast.parse(src)
    U = None
    expected = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "_U":
            U = eval(compile(ast.Expression(node.value), "seed", "eval"), {"np": np})
        if isinstance(node, ast.FunctionDef) and node.name == "test_boson_sampling_seeded":
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and getattr(sub.targets[0], "id", None) == "expected_samples":
                    expected = ast.literal_eval(sub.value)
    assert U is not None and expected is not None
    return np.array(U), [list(s) for s in expected]


def main():
    os.makedirs(OUT, exist_ok=True)
    assert refbind.available(), "build oracle/_ref first: make -C oracle ref"

    # ---- 1. the reference's own KATs
    kats = kats_from_reference_tests()
    for k in kats:  # also record what the unmodified C++ returns (double and long double)
        m = np.array(k["matrix"]["re"]).reshape(k["matrix"]["shape"]) + 1j * np.array(k["matrix"]["im"]).reshape(k["matrix"]["shape"])
        k["ref_double"] = z2l(refbind.bbfg_repeated(m, k["cols"], k["rows"], False))
        k["ref_long_double"] = z2l(refbind.bbfg_repeated(m, k["cols"], k["rows"], True))
    json.dump(kats, open(os.path.join(OUT, "kat_test_permanent.json"), "w"), indent=1)
    print("KATs:", [(k["name"], k["expected"]) for k in kats])

    # ---- 2. plain permanents, long double reference
    plain = []
    for n in list(range(1, 17)) + [18, 20, 22, 24]:
        A = unitary_group.rvs(n, random_state=n) if n > 1 else np.array([[0.6 - 0.8j]])
        plain.append({"kind": "haar_unitary", "n": n, "seed": n, "matrix": c2l(A),
                      "bbfg_long_double": z2l(refbind.bbfg(A, True)),
                      "glynn_long_double": z2l(refbind.glynn(A, True)),
                      "bbfg_double": z2l(refbind.bbfg(A, False))})
    for n in (6, 10, 14, 18, 22, 26):
        A = np.ascontiguousarray(unitary_group.rvs(2 * n, random_state=2 * n)[:n, :n])
        plain.append({"kind": "haar_subblock", "n": n, "seed": 2 * n, "matrix": c2l(A),
                      "bbfg_long_double": z2l(refbind.bbfg(A, True)),
                      "glynn_long_double": z2l(refbind.glynn(A, True)) if n <= 22 else None,
                      "bbfg_double": z2l(refbind.bbfg(A, False))})
    rng = np.random.default_rng(7)
    for shape in [(3, 5), (5, 4), (2, 2), (1, 4), (4, 3)]:
        A = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        plain.append({"kind": "non_square", "n": shape[0], "seed": 7, "matrix": c2l(A),
                      "bbfg_long_double": z2l(refbind.bbfg(A, True)), "glynn_long_double": None,
                      "bbfg_double": z2l(refbind.bbfg(A, False))})
    json.dump(plain, open(os.path.join(OUT, "plain_reference.json"), "w"))
    print("plain cases:", len(plain))

    # ---- 3. repeated-row cases (occupation patterns in the style of ctests/permanents_speed.cpp:208-231)
    rep = []
    patterns = [
        ([1, 1, 1, 1], [1, 1, 1, 1]), ([2, 2], [2, 2]), ([3, 0, 0, 1], [0, 0, 3, 1]),
        ([2, 1, 3, 0, 1, 2], [1, 1, 0, 3, 2, 2]), ([0, 0, 0, 5], [1, 1, 2, 1]),
        ([4, 4, 4], [4, 4, 4]), ([1, 0, 2, 0, 3, 0, 1], [0, 2, 0, 3, 0, 2, 0]),
        ([6, 0, 0, 0], [0, 0, 0, 6]), ([2, 2, 2, 2, 2, 2], [2, 2, 2, 2, 2, 2]),
        ([1, 2, 3, 4], [4, 3, 2, 1]), ([5, 1, 1, 1, 1, 1], [2, 2, 2, 2, 1, 1]),
        ([3, 3, 3, 3, 3], [1, 2, 3, 4, 5]), ([0, 0, 0], [0, 0, 0]), ([7], [7]),
        ([1] * 10, [1] * 10), ([2] * 8, [2] * 8),
        ([3, 1, 0, 2, 2, 0, 1, 1], [1, 1, 1, 1, 1, 1, 2, 2]),
        ([2, 0, 1, 0, 3, 1, 0, 2, 1, 0, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]),
    ]
    for idx, (rows, cols) in enumerate(patterns):
        m = len(rows)
        U = unitary_group.rvs(m, random_state=100 + idx) if m > 1 else np.array([[0.3 + 0.9j]])
        rep.append({"rows": rows, "cols": cols, "seed": 100 + idx, "matrix": c2l(U),
                    "bbfg_rep_long_double": z2l(refbind.bbfg_repeated(U, cols, rows, True)),
                    "bbfg_rep_double": z2l(refbind.bbfg_repeated(U, cols, rows, False)),
                    "glynn_rep_long_double": z2l(refbind.glynn_repeated(U, cols, rows, True)),
                    "chinhuh": z2l(refbind.chinhuh(U, cols, rows)) if sum(rows) else None})
    # rectangular compact problems (filtered states), as compute_pmf builds them
    rngs = np.random.default_rng(11)
    for idx in range(6):
        m = 9
        U = unitary_group.rvs(m, random_state=200 + idx)
        rr = rngs.multinomial(6, [1 / m] * m)
        cc = rngs.multinomial(6, [1 / m] * m)
        sub = np.ascontiguousarray(U[np.ix_(rr > 0, cc > 0)])
        rep.append({"rows": rr[rr > 0].tolist(), "cols": cc[cc > 0].tolist(), "seed": 200 + idx,
                    "matrix": c2l(sub),
                    "bbfg_rep_long_double": z2l(refbind.bbfg_repeated(sub, cc[cc > 0], rr[rr > 0], True)),
                    "bbfg_rep_double": z2l(refbind.bbfg_repeated(sub, cc[cc > 0], rr[rr > 0], False)),
                    "glynn_rep_long_double": None, "chinhuh": None})
    json.dump(rep, open(os.path.join(OUT, "repeated_reference.json"), "w"))
    print("repeated cases:", len(rep))

    # ---- 4. sampler: _U, reference python-level expectation, C++ generation order, pmf rows
    U, expected = seeding_from_reference_tests()
    cpp = refbind.simulate_B(U, 123, [2, 1, 1, 0, 1], 20)
    assert sorted(map(tuple, cpp.tolist())) == sorted(map(tuple, expected)), "reference build disagrees with its own test"
    pmfs = []
    for sample, cur in [([0, 0, 0, 0, 0], [1, 0, 0, 0, 0]), ([0, 1, 0, 0, 0], [1, 1, 0, 0, 0]),
                        ([0, 1, 0, 1, 0], [2, 1, 0, 0, 0]), ([1, 1, 0, 1, 0], [2, 1, 1, 0, 0]),
                        ([1, 1, 0, 2, 0], [2, 1, 1, 0, 1]), ([0, 0, 0, 4, 0], [2, 1, 1, 0, 1])]:
        pmfs.append({"sample": sample, "current_input": cur,
                     "pmf": refbind.compute_pmf(U, sample, cur).tolist()})
    json.dump({"U": c2l(U), "input": [2, 1, 1, 0, 1], "seed": 123, "shots": 20,
               "expected_python_order": expected, "cpp_generation_order": cpp.tolist(),
               "pmf_rows": pmfs, "source": "unittests/test_seeding.py:51-83"},
              open(os.path.join(OUT, "sampling_seed123.json"), "w"), indent=1)
    print("sampler golden ok; first shots", cpp[:3].tolist())

    # ---- 5. lossy strategies.  (a) the reference's own two LossyInterferometer KATs
    # (unittests/test_seeding.py:88-165) replayed through the call chain of
    # piquassoboost/sampling/calculations.py:105-152; (b) direct runs of the three strategies.
    exp = lossy_expected_from_reference_tests()
    inp = [2, 1, 1, 0, 1]
    lossy = {"input": inp, "seed": 123, "shots": 20, "source": "unittests/test_seeding.py:88-165"}
    # uniform: interferometer / s0 with transmissivity s0
    M = U @ np.diag([0.9] * 5) @ U @ U.T
    sv = np.linalg.svd(M)
    got = refbind.simulate_B_uniform_losses(np.ascontiguousarray(M / sv[1][0]), 123, inp, 20, float(sv[1][0]))
    want = exp["test_LossyInterferometer_boson_sampling_uniform_losses"]
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want)), "uniform-loss KAT mismatch"
    lossy["uniform"] = {"matrix": c2l(M / sv[1][0]), "transmissivity": float(sv[1][0]),
                        "cpp_generation_order": got.tolist(), "expected_python_order": want}
    # non-uniform: lossless B strategy on the 2m x 2m dilation, samples trimmed to m modes
    M = U @ np.diag([0.9, 0.8, 0.7, 0.6, 0.5]) @ U @ U.T
    E = np.ascontiguousarray(expanded_space(np.linalg.svd(M)))
    got = refbind.simulate_B(E, 123, inp + [0] * 5, 20)
    want = exp["test_LossyInterferometer_boson_sampling_seeded"]
    assert sorted(map(tuple, got[:, :5].tolist())) == sorted(map(tuple, want)), "lossy KAT mismatch"
    lossy["dilation"] = {"matrix": c2l(E), "input": inp + [0] * 5, "cpp_generation_order": got.tolist(),
                         "expected_python_order_trimmed": want}
    # direct strategy runs on seeded inputs
    runs = []
    for idx, (m, n_in, shots) in enumerate([(6, 5, 30), (8, 6, 30), (10, 8, 20)]):
        V = unitary_group.rvs(m, random_state=300 + idx)
        rng = np.random.default_rng(300 + idx)
        state = rng.multinomial(n_in - 1, [1 / m] * m)
        state[idx % 2] += 1  # the reference dereferences an empty table when the approximated modes hold no photon
        state = state.tolist()
        loss = V * rng.uniform(0.6, 0.95, size=m)[:, None]
        seed = 1000 + idx
        runs.append({
            "m": m, "input": state, "seed": seed, "shots": shots, "unitary": c2l(V), "lossy_matrix": c2l(loss),
            "uniform_losses": {"transmissivity": 0.8,
                               "samples": refbind.simulate_B_uniform_losses(V, seed, state, shots, 0.8).tolist()},
            "lossy_0": refbind.simulate_B_lossy(loss, seed, state, shots, 0).tolist(),
            "lossy_2": refbind.simulate_B_lossy(loss, seed, state, shots, 2).tolist(),
            "two_calls": refbind.simulate_B_two_calls(V, seed, state, 7, shots - 7).tolist(),
        })
    lossy["runs"] = runs
    json.dump(lossy, open(os.path.join(OUT, "sampling_lossy.json"), "w"))
    print("lossy golden ok")


if __name__ == "__main__":
    main()
