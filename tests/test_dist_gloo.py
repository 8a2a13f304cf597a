"""world_size-2 gloo tests (CPU): the bootstrap all-gather callback the native process group
calls, and rank-agreement of the shard plan.  The data path itself needs GPUs (tests/test_gpu_multi.py)."""
import ctypes
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mxnet_b200 as mx
    ag = mx.dist.make_allgather()
    # 72-byte blobs like {cudaIpcMemHandle_t, size}
    send = (ctypes.c_uint8 * 72)(*[(rank * 31 + i) % 256 for i in range(72)])
    recv = (ctypes.c_uint8 * (72 * world))()
    rc = ag(ctypes.addressof(send), 72, ctypes.addressof(recv), None)
    ok = rc == 0
    for r in range(world):
        ok = ok and list(recv[r * 72:(r + 1) * 72]) == [(r * 31 + i) % 256 for i in range(72)]
    # every rank derives the same partition of every key
    plan = [mx.dist.shard_range(s, world, r) for s in (1000, 1 << 20, 25_557_032) for r in range(world)]
    t = torch.tensor([x for p in plan for x in p], dtype=torch.int64)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    ok = ok and all(torch.equal(o, t) for o in outs)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bootstrap_allgather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _node_worker(rank, world, local_world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mxnet_b200 as mx
    node, nodes, lrank = mx.dist.node_layout(rank, world, local_world)
    local_groups, inter_groups = mx.dist.make_node_groups(world, local_world)
    # the node-local bootstrap all-gather sees the node only ...
    ag = mx.dist.make_allgather(local_groups[node])
    send = (ctypes.c_uint8 * 8)(*[rank] * 8)
    recv = (ctypes.c_uint8 * (8 * local_world))()
    ok = ag(ctypes.addressof(send), 8, ctypes.addressof(recv), None) == 0
    ok = ok and [recv[i * 8] for i in range(local_world)] == list(range(node * local_world, (node + 1) * local_world))
    # ... and the inter-node group joins the ranks of equal local rank: a sum over it is what the engine asks for
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, group=inter_groups[lrank])
    ok = ok and float(t) == sum(r + 1 for r in range(lrank, world, local_world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_node_groups_of_a_multi_node_job_world4():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_node_worker, args=(r, 4, 2, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(4)]
