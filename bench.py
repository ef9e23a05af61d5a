#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: one n x n complex Glynn permanent per step,
its 2^(n-1) Gray steps sharded over the N GPUs of one node (one process per GPU under torchrun),
partials combined with one NCCL all-gather of 4 doubles per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 40] [--impl b200|reference]

Prints ONE JSON line (rank 0).  `value` = Gray steps per second of the whole job (inputs resident:
the matrix travels as 25.6 KB of kernel parameters, there is nothing else to stage); `e2e` = the same
through the public host API (numpy in, complex out) with the host->device parameter copy and the
device->host read of the result inside the timed region -- GlynnPermanentCalculator_wrapper(...,
ndev=1) on one GPU, distributed.permanent_nccl on every rank for N > 1; `roofline` compares the
kernel's algorithmic FP64 flop rate (8n per Gray step) with the FP64 pipe peak (SMs x 64 lanes x 2 x
the SM clock sampled under load; MEASURED_PEAKS.json has no FP64 entry); `configs` carries the other
BASELINE.json configs (C1, C2, C4, C5) measured in the same run; `cpu_baseline` times the UNMODIFIED
reference BBFGPermanentCalculator (oracle/_ref, TBB shim, all host cores) on a bounded sample.
`--impl reference` runs only that CPU arm, in the same units.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "glynn_permanent_gray_steps_per_s"
UNIT = "Gray-steps/s"


def haar(n):
    from scipy.stats import unitary_group
    return np.ascontiguousarray(unitary_group.rvs(n, random_state=n), dtype=np.complex128)


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, flag in zip(names, r[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(n_target, budget_s, long_double=True):
    """Time the unmodified reference BBFGPermanentCalculator (oracle/_ref) -- or, if that library did
    not travel, the C restatement -- on a Haar unitary sized for ~budget_s seconds on this host.
    Returns (n40-equivalent Gray steps/s, descriptor dict)."""
    from oracle import oraclebind, refbind
    if refbind.available():
        kind, cores = "reference", refbind.hardware_concurrency()
        fn = lambda a: refbind.bbfg(a, long_double)
    else:
        kind, cores = "port", oraclebind.threads()
        fn = lambda a: oraclebind.bbfg(a, oraclebind.LONG_DOUBLE if long_double else oraclebind.DOUBLE)
    n = 18
    fn(haar(n))                                            # untimed: thread pool start-up, page-in
    t0 = time.perf_counter(); fn(haar(n)); dt = time.perf_counter() - t0
    rate = (2.0 ** (n - 1)) / max(dt, 1e-6) * n          # row-updates x columns per second ~ const
    while n < min(n_target, 34):
        nxt = n + 1
        if (2.0 ** (nxt - 1)) * nxt / rate > budget_s:
            break
        n = nxt
    A = haar(n)
    t0 = time.perf_counter(); val = fn(A); dt = time.perf_counter() - t0
    steps_per_s = (2.0 ** (n - 1)) / dt
    equiv = steps_per_s * n / n_target                     # a Gray step costs ~n column updates
    desc = {"value": equiv, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": (f"one n={n} Haar-unitary permanent, BBFGPermanentCalculator "
                       f"{'long double' if long_double else 'double'}, {dt:.2f} s, "
                       f"{steps_per_s:.3e} raw steps/s scaled by {n}/{n_target} to n={n_target}-equivalent steps"),
            "seconds_per_n%d_permanent_extrapolated" % n_target: (2.0 ** (n_target - 1)) / equiv}
    return equiv, desc, val


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals = []
    desc = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        v, desc, _ = cpu_reference_rate(args.n, args.ref_step_seconds)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    desc["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (2.0 ** (args.n - 1)) / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (x87 long double accumulate)",
            "data": "synthetic",
            "config": {"workload": f"n={args.n} Haar-unitary complex permanent (BASELINE config C3), CPU arm on a bounded sample"},
            "cpu_baseline": desc,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import piquassoboost_b200 as pqb
    from piquassoboost_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = pqb.load()
    n = args.n
    A_host = torch.from_numpy(haar(n)).pin_memory()        # the step's input lives in pinned host memory
    A_np = A_host.numpy()
    partial = torch.zeros(4, dtype=torch.float64, device=dev)
    gathered = torch.zeros(4 * world, dtype=torch.float64, device=dev)
    # a real (non-default) stream: its handle is passed to the C ABI so that the kernel, the NCCL
    # all-gather and the timing events are all ordered on the same stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    from piquassoboost_b200.distributed import combine_partials, gpu_shard_fn, nccl_init, nccl_shutdown

    if world > 1:
        # the library's own communicator: torch.distributed only carries rank 0's 128-byte NCCL id
        def bcast(ident):
            box = [ident]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        nccl_init(rank, world, local_rank, bcast)
    a_ptr = _lib.dptr(A_np)

    def step():
        if world > 1:
            # shard kernel + ONE ncclAllGather (4 doubles per rank), both enqueued by the C ABI on `stream`
            _lib.check(lib.pqb200_permanent_nccl_async(a_ptr, n, n, n, ctypes.c_void_p(gathered.data_ptr()),
                                                       ctypes.c_void_p(stream.cuda_stream)))
        else:
            gpu_shard_fn(A_np, rank, world, local_rank, d_out=partial.data_ptr(), stream=stream.cuda_stream)
            gathered.copy_(partial)

    def result():
        return combine_partials(gathered.cpu().numpy(), n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value_perm = result()

    # ---- kernel-only timing for the roofline: this rank's shard kernel alone, CUDA events on its stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kreps = max(1, min(args.steps, 3))
    k0.record()
    for _ in range(kreps):
        gpu_shard_fn(A_np, rank, world, local_rank, d_out=partial.data_ptr(), stream=stream.cuda_stream)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / kreps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the PUBLIC host API: numpy in -> complex out, every step.
    #      1 GPU : GlynnPermanentCalculator_wrapper(A, lib, ndev=1).calculate()  (exactly this rank's device)
    #      N GPUs: distributed.permanent_nccl(A) on every rank (pqb200_permanent_nccl: shard kernel +
    #              ncclAllGather + host fold inside the library)
    from piquassoboost_b200.distributed import permanent_nccl

    def e2e_call():
        if world == 1:
            return pqb.GlynnPermanentCalculator_wrapper(A_np, pqb.GlynnB200, ndev=1).calculate()
        return permanent_nccl(A_np)

    small = haar(16)
    if world == 1:
        pqb.GlynnPermanentCalculator_wrapper(small, pqb.GlynnB200, ndev=1).calculate()   # warm the path (slot, module)
    else:
        permanent_nccl(small)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_val = e2e_call()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    e2e_ndev = pqb.last_stats()["ndev"]
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    if rank == 0:
        steps_per_perm = 2.0 ** (n - 1)
        value = steps_per_perm * args.steps / (total_ms * 1e-3)
        # FP64 roofline.  Denominator: the FP64 pipe's peak = SMs x 64 lanes x 2 flop x the SM clock measured
        # under load during the timed region (MEASURED_PEAKS.json has no FP64 entry; the profiling guide
        # states no FP64 fallback).  The live issue-rate measurement of pqb200_fp64_peak is printed beside it.
        props = torch.cuda.get_device_properties(local_rank)
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_tflops = props.multi_processor_count * 64 * 2 * sm_mhz * 1e6 / 1e12
        dfma = ctypes.c_double(0.0)
        _lib.check(lib.pqb200_fp64_peak(local_rank, 0, 0.25, ctypes.byref(dfma)))
        dadd = ctypes.c_double(0.0)
        _lib.check(lib.pqb200_fp64_peak(local_rank, 1, 0.25, ctypes.byref(dadd)))
        shard_steps = steps_per_perm / world
        achieved = 8.0 * n * shard_steps / (kernel_ms * 1e-3) / 1e12
        instr_per_step = 5.5 * n - 6.5
        # small parity gates inside the same run (the n=40 value itself has no affordable CPU truth)
        from oracle import oraclebind, refbind

        def cpu_truth(M):
            return refbind.bbfg(M, True) if refbind.available() else oraclebind.bbfg(M, oraclebind.LONG_DOUBLE)

        parity = []
        for pn in (18, 24):
            B = haar(pn)
            got = pqb.GlynnPermanentCalculator_wrapper(B, pqb.GlynnB200, ndev=1).calculate()
            want = cpu_truth(B)
            parity.append({"n": pn, "rel_err_vs_long_double_reference": abs(got - want) / abs(want), "tolerance": 1e-10})
        cpu_equiv, cpu_desc, _ = cpu_reference_rate(n, args.cpu_seconds)
        # the value of THIS run's permanent against the committed golden values: the reference's own long
        # double where one exists (n <= 38, tools/ref_big.py), else the double-double evaluation (itself
        # pinned to the reference's long double up to n=40; tests/golden/plain_dd_reference.json)
        accuracy = None
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plain_dd_reference.json")))["cases"]
            if str(n) in gold:
                g = complex(*[float.fromhex(x) for x in gold[str(n)]["dd_parts_hex"]][0::2])
                accuracy = {"rel_err_vs_double_double_golden": abs(value_perm - g) / abs(g),
                            "tolerance": 1e-10, "golden": "tests/golden/plain_dd_reference.json"}
        except OSError:
            pass
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (double-double accumulation)",
            "data": "synthetic",
            "config": {"workload": f"n={n} Haar-unitary (scipy unitary_group, seed {n}) complex permanent, "
                                   f"2^{n-1} Gray steps per step sharded over {world} GPU(s)"
                                   + (", one ncclAllGather of 4 doubles per rank issued by the C ABI (pqb200_permanent_nccl_async) "
                                      "on the kernel's stream" if world > 1 else ""),
                       "l2_note": "working set is 25.6 KB of kernel parameters + registers; no HBM traffic to flush"},
            "seconds_per_permanent": total_ms / args.steps * 1e-3,
            "permanent": [value_perm.real, value_perm.imag],
            "parity_check": parity,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": f"{props.multi_processor_count} SMs x 64 FP64 lanes x 2 flop x {sm_mhz:.0f} MHz "
                                        "(median SM clock sampled during the timed region)",
                         "measured_issue_rates": {
                             "dfma_tflops": 2.0 * dfma.value / 1e12, "dadd_thread_instr_per_s": dadd.value,
                             "note": "pqb200_fp64_peak: independent DFMA / DADD chains, 2 CTAs x 256 threads per SM"},
                         "algorithmic_flops_per_step": 8 * n, "kernel_ms": kernel_ms,
                         # FP64 instructions per Gray step after the real-row/real-column normalisation:
                         # 5.5n - 6.5; each issues at most 1 FMA = 2 flop
                         "fp64_instr_per_step": instr_per_step,
                         "fp64_pipe_utilisation": instr_per_step * shard_steps / (kernel_ms * 1e-3)
                                                  / (props.multi_processor_count * 64 * sm_mhz * 1e6),
                         "instruction_mix_ceiling": (8.0 * n) / (2.0 * instr_per_step)},
            "accuracy": accuracy,
            "cpu_baseline": cpu_desc,
            "e2e": {"value": steps_per_perm * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": 16 * n * n * world, "d2h_bytes_per_step": 32 * world,
                    "seconds_per_permanent": e2e_s / args.steps,
                    "api": ("GlynnPermanentCalculator_wrapper(A, GlynnB200, ndev=1).calculate()" if world == 1
                            else "distributed.permanent_nccl(A) on every rank"),
                    "devices_used_per_process": e2e_ndev,
                    "permanent": [e2e_val.real, e2e_val.imag]},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
        }
        if not args.no_configs:
            line["configs"] = baseline_configs(pqb, world, cpu_truth, props.multi_processor_count * 64 * sm_mhz * 1e6)
        print(json.dumps(line), flush=True)
    if world > 1:
        # the other ranks wait on the CPU (a NCCL barrier would park a spinning kernel on their GPUs
        # while rank 0 runs the C5 side measurement over all devices of the box)
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("pqb200_bench_done", "1")
        else:
            store.wait(["pqb200_bench_done"], datetime.timedelta(minutes=30))
        nccl_shutdown()
        dist.destroy_process_group()


def ncu_traffic():
    """dram bytes (read + write) per launch of the headline kernel from the committed ncu --set full
    capture of this round's kernel (profiles/); (None, reason) when no capture is committed."""
    import csv
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_plain40_r*_raw.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            hdr = rows[0]
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            units, data = rows[1], [r for r in rows[2:] if len(r) > max(ir, iw)]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = [float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0)
                   + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0) for r in data]
            if tot:
                return sum(tot) / len(tot), f"ncu --set full capture {os.path.relpath(path, ROOT)} (dram read + write per launch)"
        except (OSError, ValueError, IndexError):
            continue
    return None, "no ncu capture committed under profiles/"


def baseline_configs(pqb, world, cpu_truth, fp64_instr_peak):
    """The other BASELINE.json configs in the same run (rank 0, after the timed region): C1 n=20, C2 n=30,
    C4 30 photons / 60 modes, C5 20 photons / 40 modes / 1000 shots -- through the public classes, wall
    clock per call (host buffers in and out), kernel time from the library's CUDA events, parity against
    the CPU reference where that takes seconds.  C5 shards its shots over all `world` GPUs of the box from
    this one process (the `ndev` route; no collective)."""
    from scipy.stats import unitary_group
    from oracle import oraclebind, refbind
    out = {}

    def timed(fn, reps):
        fn()
        t = time.perf_counter()
        for _ in range(reps):
            v = fn()
        return v, (time.perf_counter() - t) / reps

    for tag, pn in (("C1_n20", 20), ("C2_n30", 30)):
        M = haar(pn)
        v, sec = timed(lambda: pqb.GlynnPermanentCalculator_wrapper(M, pqb.GlynnB200, ndev=1).calculate(), 5)
        st = pqb.last_stats()
        entry = {"seconds_per_call": sec, "kernel_ms": st["kernel_ms"], "gray_steps_per_s": st["gray_steps"] / sec,
                 "fp64_pipe_utilisation_kernel": (5.5 * pn - 6.5) * st["gray_steps"] / (st["kernel_ms"] * 1e-3) / fp64_instr_peak}
        if pn <= 26 or refbind.available():
            t = time.perf_counter()
            want = cpu_truth(M)
            entry["reference_long_double_s"] = time.perf_counter() - t
            entry["rel_err"] = abs(v - want) / abs(want)
        out[tag] = entry
    U = np.ascontiguousarray(unitary_group.rvs(60, random_state=4))
    rng = np.random.default_rng(4)
    inp = np.array([1] * 30 + [0] * 30)
    cases = {"15x2": np.array([2] * 15 + [0] * 45), "10x3": np.array([3] * 10 + [0] * 50),
             "multinomial": rng.multinomial(30, rng.dirichlet(np.ones(60) * 0.3))}
    for name, o in cases.items():
        v, sec = timed(lambda: pqb.ChinHuhPermanentCalculator_wrapper(pqb.GlynnRepB200, U, inp, o, ndev=1).calculate(), 5)
        st = pqb.last_stats()
        entry = {"seconds_per_call": sec, "kernel_ms": st["kernel_ms"], "gray_steps": st["gray_steps"]}
        if refbind.available():
            t = time.perf_counter()
            want = refbind.bbfg_repeated(U, inp, o, True)
            entry["reference_long_double_s"] = time.perf_counter() - t
            entry["rel_err"] = abs(v - want) / abs(want)
        out["C4_30ph_60modes_" + name] = entry
    V = np.ascontiguousarray(unitary_group.rvs(40, random_state=5))
    state = np.array([1] * 20 + [0] * 20, dtype=np.int64)
    for shots in (1000, 20000):
        for nd in sorted({1, world}):
            sim = pqb.GeneralizedCliffordsBSimulationStrategy(V, seed=5, ndev=nd)
            sim.simulate(state, shots)       # first call on these devices: contexts, buffers
            sim.seed(5)
            t = time.perf_counter()
            res = sim.simulate(state, shots)
            sec = time.perf_counter() - t
            st = pqb.last_stats()
            entry = {"shots": shots, "gpus": st["ndev"], "seconds_per_call": sec, "shots_per_s": shots / sec,
                     "kernel_ms_device0": st["kernel_ms"], "gray_terms": st["gray_steps"], "launches": st["launches"]}
            if shots == 1000 and refbind.available():
                t = time.perf_counter()
                want = refbind.simulate_B(V, 5, state, 20)
                entry["reference_s_per_shot"] = (time.perf_counter() - t) / 20
                entry["first_20_shots_identical_to_reference"] = bool(np.array_equal(np.array(res[:20]), want))
            out[f"C5_20ph_40modes_{shots}shots_{nd}gpu"] = entry
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=40, help="matrix size (BASELINE config C3 = 40)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--ref-step-seconds", type=float, default=6.0, help="--impl reference: per-step budget")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C2/C4/C5 side measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
