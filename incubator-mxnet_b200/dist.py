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
_ALLREDUCE_PROTO = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p)
_state = {"cb": None, "group": None, "rank": 0, "world": 1, "dev": None, "ar": None}


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


def node_layout(rank, world, local_world):
    """(node_rank, num_nodes, local_rank) of ``rank`` when ranks are numbered node by node, ``local_world`` to a node
    (what torchrun does)."""
    assert local_world >= 1 and world % local_world == 0, "world size %d is not a multiple of %d ranks per node" % (
        world, local_world)
    return rank // local_world, world // local_world, rank % local_world


def make_node_groups(world, local_world, backend=None):
    """torch.distributed sub-groups of a multi-node job: one per node (its ranks) and one per local rank (the ranks
    holding it on every node).  Every rank creates every group, in the same order (torch.distributed's rule)."""
    import torch.distributed as dist
    nodes = world // local_world
    local_groups = [dist.new_group(list(range(i * local_world, (i + 1) * local_world)), backend=backend)
                    for i in range(nodes)]
    inter_groups = [dist.new_group(list(range(j, world, local_world)), backend=backend) for j in range(local_world)]
    return local_groups, inter_groups


def set_hierarchy(node_rank, num_nodes, allreduce):
    """Describe the nodes of a multi-node job to the engine (MXKVB200SetHierarchy); the group given to
    ``init_process_group`` / ``init_with_allgather`` before is then ONE node.  ``allreduce(dev_ptr, count, dtype,
    cuda_stream)`` sums ``count`` elements of mshadow type ``dtype`` at device address ``dev_ptr`` in place over the
    ranks with this rank's local rank on every node, ordered on ``cuda_stream``.  Afterwards
    ``mx.kv.create('dist_device_sync')`` works."""
    def _cb(ptr, count, dtype, stream, _ctx):
        try:
            allreduce(ptr, count, dtype, stream)
            return 0
        except Exception as e:  # never raise through the C boundary
            import sys
            sys.stderr.write("mxnet_b200.dist inter-node all-reduce failed: %r\n" % (e,))
            return 1

    cb = _ALLREDUCE_PROTO(_cb)
    check_call(_LIB.MXKVB200SetHierarchy(node_rank, num_nodes, cb, None))
    _state.update(ar=cb)
    return node_rank, num_nodes


def make_torch_allreduce(inter_group, device):
    """The inter-node sum on top of torch.distributed (NCCL): the collective is enqueued behind the engine's
    stream and the engine's stream waits for it -- nothing blocks the host."""
    import torch
    import torch.distributed as dist
    # mshadow type flags (3rdparty/mshadow/mshadow/base.h:352-366) -> (typestr of the raw view, final torch dtype)
    kinds = {0: ("<f4", None), 1: ("<f8", None), 2: ("<f2", None), 3: ("|u1", None), 4: ("<i4", None),
             5: ("|i1", None), 6: ("<i8", None), 12: ("<i2", torch.bfloat16)}

    class _Raw(object):
        def __init__(self, ptr, count, typestr):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}

    def _allreduce(ptr, count, dtype, stream):
        typestr, view = kinds[dtype]
        t = torch.as_tensor(_Raw(ptr, count, typestr), device=torch.device("cuda", device))
        if view is not None:
            t = t.view(view)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=torch.device("cuda", device))):
            dist.all_reduce(t, group=inter_group, async_op=True).wait()

    return _allreduce


def init_process_group(device=None, group=None, local_world=None):
    """Bind the engine to the already-initialised torch.distributed group: rank r drives GPU
    ``device`` (default LOCAL_RANK).  Must run before the first use of that GPU by the engine.

    Multi-node jobs pass ``local_world`` (ranks per node; default ``MXKV_B200_LOCAL_WORLD``, else
    ``LOCAL_WORLD_SIZE`` as set by torchrun): the engine's peer-memory group is then the node, the nodes are joined
    by NCCL all-reduces between ranks of equal local rank, and ``mx.kv.create('dist_device_sync')`` gives the
    hierarchical store.  (Setting it below the real node size splits one box into several "nodes".)"""
    import torch
    import torch.distributed as dist
    assert dist.is_initialized(), "call torch.distributed.init_process_group first"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(device)
    if local_world is None:
        local_world = int(os.environ.get("MXKV_B200_LOCAL_WORLD", os.environ.get("LOCAL_WORLD_SIZE", world)))
    if group is None and 0 < local_world < world:
        node, nodes, lrank = node_layout(rank, world, local_world)
        local_groups, inter_groups = make_node_groups(world, local_world)
        cb = _ALLGATHER_PROTO(make_allgather(local_groups[node], device))
        check_call(_LIB.MXKVB200CommInit(lrank, local_world, device, cb, None))
        _state.update(cb=cb, group=local_groups[node], rank=lrank, world=local_world, dev=device,
                      inter_group=inter_groups[lrank], all_groups=(local_groups, inter_groups))
        set_hierarchy(node, nodes, make_torch_allreduce(inter_groups[lrank], device))
        return rank, world
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
