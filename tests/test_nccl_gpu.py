"""The collective route that lives inside the library (pqb200_nccl_* / pqb200_permanent_nccl): a
1-rank communicator on one GPU, and 2 ranks (one process per GPU) when two GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import haar_unitary, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_communicator(ref):
    import piquassoboost_b200 as pqb
    from piquassoboost_b200.distributed import nccl_init, nccl_shutdown, permanent_nccl
    pqb.load()
    nccl_init(0, 1, 0, lambda ident: ident)
    try:
        A = haar_unitary(16, 16)
        got = permanent_nccl(A)
        assert relerr(got, ref.bbfg(A, True)) < 1e-11
        assert got == pqb.GlynnPermanent(A, ).calculate() or relerr(got, pqb.GlynnPermanent(A).calculate()) < 1e-14
        with pytest.raises(Exception):
            nccl_init(0, 1, 0, lambda ident: ident)          # second init without shutdown
    finally:
        nccl_shutdown()
    with pytest.raises(Exception):
        permanent_nccl(haar_unitary(6, 6))                   # no communicator


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from piquassoboost_b200.distributed import nccl_init, nccl_shutdown, permanent_nccl
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def bcast(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    nccl_init(rank, world, rank, bcast)
    vals = [permanent_nccl(haar_unitary(n, n)) for n in (3, 12, 22)]
    q.put((rank, vals))
    nccl_shutdown()
    dist.destroy_process_group()


def test_two_ranks_two_gpus(ref):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == results[1]                          # identical bits on every rank
    for n, v in zip((3, 12, 22), results[0]):
        assert relerr(v, ref.bbfg(haar_unitary(n, n), True)) < 1e-11
