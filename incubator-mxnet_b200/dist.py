"""One-process-per-GPU bootstrap over ``torch.distributed`` (plumbing only).

The native engine needs exactly one host collective -- an all-gather of small byte blobs --
to exchange CUDA IPC handles of its peer-mapped arena and to check that collective
allocations agree.  ``init_process_group`` wires that to ``torch.distributed`` (NCCL on a GPU
box, gloo on CPU for tests).  Nothing on the data path goes through torch.distributed: the
gradient exchange is the engine's own kernel reading/writing peer memory over NVLink.
"""
import ctypes
import os

from .base import _LIB, check_call

_ALLGATHER_PROTO = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_void_p)
_state = {"cb": None, "group": None, "rank": 0, "world": 1, "dev": None}


def make_allgather(group=None, device=None):
    """Python implementation of the MXKVB200AllGatherFn contract on top of torch.distributed.
    Returns a callable (send_ptr, nbytes, recv_ptr, ctx) -> int.  Usable (and tested) on CPU/gloo."""
    import torch
    import torch.distributed as dist

    def _allgather(send, nbytes, recv, _ctx):
        try:
            world = dist.get_world_size(group)
            raw = ctypes.string_at(send, nbytes)
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            backend = dist.get_backend(group)
            if backend == "nccl":
                t = t.cuda(device)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t, group=group)
            for i, o in enumerate(outs):
                b = o.cpu().numpy().tobytes()
                ctypes.memmove(recv + i * nbytes, b, nbytes)
            return 0
        except Exception as e:  # never raise through the C boundary
            import sys
            sys.stderr.write("mxnet_b200.dist all-gather failed: %r\n" % (e,))
            return 1

    return _allgather


def init_process_group(device=None, group=None):
    """Bind the engine to the already-initialised torch.distributed group: rank r drives GPU
    ``device`` (default LOCAL_RANK).  Must run before the first use of that GPU by the engine."""
    import torch
    import torch.distributed as dist
    assert dist.is_initialized(), "call torch.distributed.init_process_group first"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(device)
    cb = _ALLGATHER_PROTO(make_allgather(group, device))
    check_call(_LIB.MXKVB200CommInit(rank, world, device, cb, None))
    _state.update(cb=cb, group=group, rank=rank, world=world, dev=device)
    return rank, world


def init_with_allgather(rank, world, device, allgather):
    """Bind the engine to ANY process group: all it needs is ``allgather(payload: bytes) -> [bytes] * world``
    (used once per collective allocation, to exchange CUDA IPC handles).  ``init_process_group`` is this with
    torch.distributed behind it; MPI, a TCP store, or -- in the simulator tests -- files work as well."""
    def _cb(send, nbytes, recv, _ctx):
        try:
            parts = allgather(ctypes.string_at(send, nbytes))
            assert len(parts) == world
            for i, b in enumerate(parts):
                ctypes.memmove(recv + i * nbytes, bytes(b), nbytes)
            return 0
        except Exception as e:  # never raise through the C boundary
            import sys
            sys.stderr.write("mxnet_b200.dist all-gather failed: %r\n" % (e,))
            return 1

    cb = _ALLGATHER_PROTO(_cb)
    check_call(_LIB.MXKVB200CommInit(rank, world, device, cb, None))
    _state.update(cb=cb, group=None, rank=rank, world=world, dev=device)
    return rank, world


def destroy_process_group():
    check_call(_LIB.MXKVB200CommDestroy())
    _state.update(cb=None, group=None, rank=0, world=1, dev=None)


def shard_range(size, world, rank):
    """[begin, end) of the elements rank `rank` reduces/updates for a key of `size` elements in
    the two-shot path (the same function the native engine uses)."""
    b, e = ctypes.c_int64(), ctypes.c_int64()
    check_call(_LIB.MXKVB200ShardRange(ctypes.c_int64(size), world, rank, ctypes.byref(b), ctypes.byref(e)))
    return b.value, e.value
