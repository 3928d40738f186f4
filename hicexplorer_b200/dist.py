"""Multi-GPU plumbing: one process per GPU, nnz-balanced contiguous row blocks, and the ONE
collective of the path (a sum-allreduce of the marginal / search vector per SpMV) delegated to
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The reference has nothing distributed; this is the sharding the north star asks
for.  Each rank streams its own rows of the matrix, writes its slice of the vector
into a zeroed full-length buffer, the buffers are summed across ranks (x + 0 is exact, so this
equals an all-gather bit for bit) and the O(n) vector update then runs identically on every rank.
"""
import ctypes

import numpy as np

from . import _lib


def nnz_balanced_partition(indptr, world):
    """Row boundaries [b_0=0, ..., b_world=n]: block k ends at the first row whose cumulative entry
    count reaches ceil(k * nnz / world), so every rank streams an equal share of the 12 B/nnz."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    nnz = int(indptr[-1] - indptr[0])
    bounds = [0]
    for k in range(1, world):
        target = indptr[0] + (k * nnz + world - 1) // world
        b = int(np.searchsorted(indptr, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def lpt_assignment(weights, world):
    """Longest-processing-time assignment of independent chromosomes (``--perchr``) to ranks: owner[c] = rank.
    Deterministic, so every rank computes the same table without talking to the others."""
    loads = [0.0] * world
    owner = [0] * len(weights)
    for c in sorted(range(len(weights)), key=lambda i: (-weights[i], i)):
        r = min(range(world), key=lambda k: (loads[k], k))
        owner[c] = r
        loads[r] += weights[c]
    return owner


class _DevicePtr(object):
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def make_allreduce_callback(group=None, device_buffers=True):
    """Returns (ctypes callback, keepalive).  The callback matches ``hb_allreduce_fn`` in
    include/hicbalance.h: sum-allreduce ``count`` doubles at ``d_buf`` over the ranks, ordered on
    ``stream``.  ``device_buffers=False`` treats the pointer as host memory (gloo CPU tests)."""
    import torch
    import torch.distributed as dist

    cache = {}

    def _tensor(ptr, count):
        key = (ptr, count)
        t = cache.get(key)
        if t is None:
            if device_buffers:
                t = torch.as_tensor(_DevicePtr(ptr, count), device="cuda")
            else:
                arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(count,))
                t = torch.from_numpy(arr)
            cache[key] = t
        return t

    def _cb(_ctx, d_buf, count, stream):
        try:
            t = _tensor(d_buf, count)
            if device_buffers and stream:
                with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as exc:  # never let an exception cross the C boundary
            import sys
            sys.stderr.write("allreduce callback failed: %r\n" % (exc,))
            return 1

    fn = _lib.ALLREDUCE_FN(_cb)
    return fn, (fn, cache)


def attach_comm(dev, rank, world, group=None):
    """Register the collective on a DeviceCSR row block; returns an object to keep alive."""
    fn, keep = make_allreduce_callback(group)
    _lib.check(_lib.lib().hb_set_comm(dev.handle, int(rank), int(world), ctypes.cast(fn, ctypes.c_void_p), None))
    return keep


def attach_peer(dev, rank, world, group=None):
    """Switch a DeviceCSR row block to the peer-store exchange: the SpMV epilogue writes each row straight into every
    rank's replica of the exchange vector over NVLink (CUDA IPC mappings), no collective in the loop.  Needs one
    process per GPU inside one NVLink/P2P domain and world <= 8.  Returns True on success; on failure (no peer access,
    IPC unavailable) the handle is left untouched and the caller should fall back to attach_comm()."""
    import torch
    import torch.distributed as dist
    L = _lib.lib()
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    handle = (ctypes.c_ubyte * 64)()
    rc = L.hb_peer_export(dev.handle, int(rank), int(world), ctypes.cast(handle, ctypes.c_void_p))
    if rc != 0:
        ok[0] = 0
    mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
    gathered = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        blob = torch.cat(gathered).cpu().numpy().tobytes()
        rc = L.hb_peer_attach(dev.handle, blob)
        if rc != 0:
            ok[0] = 0
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also the barrier the protocol needs after attaching
    if int(ok.item()) != 1:
        # leave peer mode everywhere so that all ranks take the same fallback
        return False
    return True


def attach_peer_inprocess(devs):
    """Peer-store exchange between row-block handles that live in ONE process (``devs`` in rank order; one host thread
    per rank drives its handle afterwards): several GPUs with peer access, or several blocks on one GPU (tests)."""
    L = _lib.lib()
    world = len(devs)
    scratch = (ctypes.c_ubyte * 64)()
    for rank, dev in enumerate(devs):
        _lib.check(L.hb_peer_export(dev.handle, rank, world, ctypes.cast(scratch, ctypes.c_void_p)))
    table = (ctypes.c_void_p * world)(*[d.handle for d in devs])
    for dev in devs:
        _lib.check(L.hb_peer_attach_inprocess(dev.handle, ctypes.cast(table, ctypes.c_void_p)))
