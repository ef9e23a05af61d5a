"""One-process-per-GPU evaluation of a single permanent: every rank computes its shard of the Gray
index space, the 4-double partials are all-gathered and folded in rank order with double-double adds.
Two routes: `nccl_init` + `permanent_nccl` keep the collective INSIDE the library (its own NCCL
communicator, kernel and ncclAllGather on one stream: pqb200_permanent_nccl); `permanent_distributed`
uses the caller's torch.distributed group instead (NCCL on GPUs; gloo in the CPU tests) around
pqb200_permanent_shard / pqb200_permanent_combine.

The reference has no counterpart for splitting ONE permanent (its optional MPI mode shards shots,
CGeneralizedCliffordsBSimulationStrategy.cpp:158-231); the partition mirrors what its TBB chunking
does inside one process (BBFGPermanentCalculator.hpp:120-134).
"""
import ctypes

import numpy as np

from . import _lib


def combine_partials(parts, rows):
    """Fold `len(parts)//4` partials (re.hi re.lo im.hi im.lo each), apply 2^-(rows-1)."""
    parts = np.ascontiguousarray(parts, dtype=np.float64).reshape(-1)
    out = np.zeros(2)
    _lib.check(_lib.load().pqb200_permanent_combine(_lib.dptr(parts), parts.size // 4, rows,
                                                    _lib.dptr(out)))
    return complex(out[0], out[1])


def gpu_shard_fn(a, rank, world, device, d_out=None, stream=None):
    """This rank's partial on CUDA device `device`.  With d_out (a device pointer) the 4 doubles stay
    on the device and nothing synchronises; otherwise they are returned as a numpy array."""
    a = _lib.as_matrix(a)
    n, m = a.shape
    host = None if d_out is not None else np.zeros(4)
    _lib.check(_lib.load().pqb200_permanent_shard(
        _lib.dptr(a), n, m, a.strides[0] // 16, device, rank, world,
        None if host is None else _lib.dptr(host),
        None if d_out is None else ctypes.c_void_p(d_out),
        None if stream is None else ctypes.c_void_p(stream)))
    return host


def _point_at_bundled_nccl():
    """If the pip-bundled NCCL (the one torch links) is installed, make the library dlopen THAT file:
    two NCCL builds with the same soname cannot live in one process, whichever is loaded first."""
    import importlib.util
    import os
    if os.environ.get("PQB200_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["PQB200_NCCL_LIB"] = cand
            return


def nccl_init(rank, world, device, broadcast):
    """Create the library's own NCCL communicator (pqb200_nccl_init).  `broadcast(bytes_or_None) ->
    bytes` must hand rank 0's 128-byte id to every rank (MPI_Bcast, torch.distributed, a file ...)."""
    lib = _lib.load()
    _point_at_bundled_nccl()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.pqb200_nccl_unique_id(buf))
    ident = broadcast(buf.raw if rank == 0 else None)
    _lib.check(lib.pqb200_nccl_init(device, rank, world, ctypes.create_string_buffer(bytes(ident), 128)))


def nccl_shutdown():
    _lib.check(_lib.load().pqb200_nccl_shutdown())


def permanent_nccl(a):
    """perm(a), collectively over the ranks of the library's communicator: shard kernel + ONE
    ncclAllGather of 4 doubles per rank on the same stream, double-double fold on every rank."""
    a = _lib.as_matrix(a)
    out = np.zeros(2)
    _lib.check(_lib.load().pqb200_permanent_nccl(_lib.dptr(a), a.shape[0], a.shape[1],
                                                 a.strides[0] // 16, _lib.dptr(out)))
    return complex(out[0], out[1])


def permanent_distributed(a, group=None, device=None, shard_fn=None):
    """perm(a), computed cooperatively by all ranks of `group`; every rank returns the value.

    shard_fn(a, rank, world) -> 4 doubles lets the tests substitute the per-rank compute; the default
    is the CUDA engine on `device` (default: the current torch CUDA device)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a = _lib.as_matrix(a)
    backend = dist.get_backend(group)
    if shard_fn is None:
        if device is None:
            device = torch.cuda.current_device()
        dev = torch.device("cuda", device)
        partial = torch.zeros(4, dtype=torch.float64, device=dev)
        gathered = torch.zeros(4 * world, dtype=torch.float64, device=dev)
        # the kernel is enqueued on torch's CURRENT stream of the device -- the stream the all-gather
        # below is ordered against.  Handle 0 is the legacy default stream: cudaStreamLegacy (0x1) names
        # it explicitly (a NULL stream argument also means "the caller's legacy default stream" when
        # d_out is given, include/pqb200.h).
        handle = torch.cuda.current_stream(dev).cuda_stream
        gpu_shard_fn(a, rank, world, device, d_out=partial.data_ptr(), stream=handle or 0x1)
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        partial = torch.as_tensor(np.asarray(shard_fn(a, rank, world), dtype=np.float64)).to(dev)
        gathered = torch.zeros(4 * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, partial, group=group)
    return combine_partials(gathered.cpu().numpy(), a.shape[0])
