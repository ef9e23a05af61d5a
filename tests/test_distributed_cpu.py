"""world_size-2 gloo test of the N>1 host path (piquassoboost_b200/distributed.py): process-group
plumbing, gather order and the double-double combine + scaling of the C ABI, with the per-rank
compute substituted by a CPU shard (there is no GPU in the dev container; on the GPU box the same
function runs with the CUDA shard and NCCL in bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_shard(a, rank, world):
    """Partial Glynn sum over the Gray indices [rank*T/world, (rank+1)*T/world), long double."""
    n = a.shape[0]
    T = 1 << (n - 1)
    lo, hi = rank * T // world, (rank + 1) * T // world
    acc = np.clongdouble(0)
    al = a.astype(np.clongdouble)
    for g in range(lo, hi):
        gray = g ^ (g >> 1)
        delta = np.array([1] + [-1 if (gray >> (n - 1 - i)) & 1 else 1 for i in range(1, n)], dtype=np.longdouble)
        acc += np.prod(delta) * np.prod(delta @ al)
    hi_re, hi_im = float(acc.real), float(acc.imag)
    return [hi_re, float(acc.real - np.longdouble(hi_re)), hi_im, float(acc.imag - np.longdouble(hi_im))]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from piquassoboost_b200.distributed import permanent_distributed
    from scipy.stats import unitary_group
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    A = unitary_group.rvs(9, random_state=9)
    val = permanent_distributed(A, shard_fn=_cpu_shard)
    q.put((rank, val))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_permanent():
    import torch.multiprocessing as mp
    from scipy.stats import unitary_group
    sys.path.insert(0, ROOT)
    from oracle import oraclebind

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oraclebind.bbfg(unitary_group.rvs(9, random_state=9), oraclebind.LONG_DOUBLE)
    assert set(results) == {0, 1}
    for r in (0, 1):
        assert abs(results[r] - want) / abs(want) < 1e-13
    assert results[0] == results[1]            # every rank folds the same partials in the same order
