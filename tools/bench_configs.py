"""tools/bench_configs.py -- times and parity-checks the five BASELINE.json configs on the GPU box
(C3, the n=40 headline, is bench.py's job and is only run here with --with-c3).  Prints one JSON
line per case; the run committed under profiles/ is the evidence quoted in DESIGN.md."""
import json
import os
import random
import sys
import time

import numpy as np
from scipy.stats import unitary_group

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piquassoboost_b200 as pqb  # noqa: E402
from oracle import refbind  # noqa: E402

HAVE_REF = refbind.available()


def rel(a, b):
    return abs(a - b) / abs(b)


def timed(fn, reps=3):
    best, val = 1e30, None
    for _ in range(reps):
        t = time.perf_counter()
        val = fn()
        best = min(best, time.perf_counter() - t)
    return best, val


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    pqb.load()
    pqb.GlynnPermanent(np.eye(4)).calculate()  # context creation outside the timings
    # ---- C1: n=20 Haar unitary (permanent_benchmark.py:56-60)
    A = unitary_group.rvs(20, random_state=20)
    t, v = timed(lambda: pqb.GlynnPermanent(A).calculate(), 10)
    st = pqb.last_stats()
    tr, r = timed(lambda: refbind.bbfg(A, True), 3) if HAVE_REF else (None, None)
    emit(config="C1", n=20, gpu_wall_s=t, gpu_kernel_ms=st["kernel_ms"], ref_long_double_s=tr,
         rel_err=rel(v, r) if r else None, cores=refbind.hardware_concurrency() if HAVE_REF else None)
    # ---- C2: n=30 Haar unitary and Haar sub-block
    for name, M in (("haar_unitary", unitary_group.rvs(30, random_state=30)),
                    ("haar_subblock_of_60", np.ascontiguousarray(unitary_group.rvs(60, random_state=60)[:30, :30]))):
        t, v = timed(lambda: pqb.BBFGPermanentLongDouble(M).calculate(), 3)
        st = pqb.last_stats()
        tr, r = timed(lambda: refbind.bbfg(M, True), 1) if HAVE_REF else (None, None)
        trd, rd = timed(lambda: refbind.bbfg(M, False), 1) if HAVE_REF else (None, None)
        emit(config="C2", case=name, n=30, gpu_wall_s=t, gpu_kernel_ms=st["kernel_ms"],
             steps_per_s=2.0 ** 29 / (st["kernel_ms"] * 1e-3), ref_long_double_s=tr, ref_double_s=trd,
             rel_err_vs_long_double=rel(v, r) if r else None,
             ref_double_rel_err_vs_long_double=rel(rd, r) if r else None)
    # ---- C4: 30 photons in 60 modes with multiplicities
    U = unitary_group.rvs(60, random_state=4)
    inp = np.array([1] * 30 + [0] * 30, dtype=np.int64)
    outs = {
        "30_distinct": np.array([1] * 30 + [0] * 30),
        "15x2": np.array([2] * 15 + [0] * 45),
        "10x3": np.array([3] * 10 + [0] * 50),
        "multinomial": np.random.default_rng(4).multinomial(30, [1 / 60] * 60),
    }
    for name, out in outs.items():
        calc = pqb.BBFGRepeatedPermanentCalculatorLongDouble(U, inp, out.astype(np.int64))
        t, v = timed(calc.calculate, 3)
        st = pqb.last_stats()
        tr, r = (None, None)
        if HAVE_REF and (name != "30_distinct" or "--full" in sys.argv):
            tr, r = timed(lambda: refbind.bbfg_repeated(U, inp, out, True), 1)
        emit(config="C4", case=name, gray_steps=st["gray_steps"], gpu_wall_s=t, gpu_kernel_ms=st["kernel_ms"],
             ref_long_double_s=tr, rel_err=rel(v, r) if r else None, value=[v.real, v.imag])
    # ---- C5: Clifford-Clifford B sampling, 20 photons / 40 modes / 1000 shots, seed 123
    U = unitary_group.rvs(40, random_state=5)
    modes = random.Random(5).sample(range(40), 20)
    inp = np.zeros(40, dtype=np.int64)
    inp[modes] = 1
    shots = 1000 if "--quick" not in sys.argv else 100
    # first call: context/buffer set-up of the sampler (reported separately); second call: steady state
    t0 = time.perf_counter()
    pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=123).simulate(inp, shots)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = np.array(pqb.GeneralizedCliffordsBSimulationStrategy(U, seed=123).simulate(inp, shots))
    t = time.perf_counter() - t0
    st = pqb.last_stats()
    tr, same = None, None
    if HAVE_REF:
        nref = 20
        t0 = time.perf_counter()
        want = refbind.simulate_B(U, 123, inp, nref)
        tr = (time.perf_counter() - t0) / nref
        same = bool((got[:nref] == want).all())
    emit(config="C5", shots=shots, gpu_wall_s=t, gpu_first_call_wall_s=t_first, gpu_kernel_ms_sum=st["kernel_ms"], launches=st["launches"],
         gpu_s_per_shot=t / shots, ref_s_per_shot=tr, first_20_shots_identical=same,
         mean_occupied_modes=float((got > 0).sum(axis=1).mean()))
    if "--with-c3" in sys.argv:
        A = unitary_group.rvs(40, random_state=40)
        t, v = timed(lambda: pqb.GlynnPermanent(A).calculate(), 1)
        emit(config="C3", n=40, gpu_wall_s=t, ndev=pqb.last_stats()["ndev"], value=[v.real, v.imag])


if __name__ == "__main__":
    main()
