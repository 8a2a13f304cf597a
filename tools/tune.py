"""Sweep the scheduling knobs of the dense kernel on the bench workload (run under gpurun).
python tools/tune.py            # N=1
torchrun ... tools/tune.py      # N>1
Prints one line per configuration: kernel ms (max over ranks) and the implied bandwidth."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mxnet_b200 as mx                      # noqa: E402
from mxnet_b200.base import _LIB, check_call  # noqa: E402
from bench import keyset, nelem              # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        mx.dist.init_process_group(device=local)
    shapes = keyset(os.environ.get("TUNE_WORKLOAD", "sweep"))
    S = 4 * sum(nelem(s) for s in shapes)
    keys = list(range(len(shapes)))
    rng = np.random.default_rng(rank)
    nval = int(os.environ.get("TUNE_NVAL", 1))      # values per key on this GPU (emulates n sources)
    alloc = mx.nd.empty_multicast if os.environ.get("TUNE_ALLOC", "symmetric") == "multicast" else mx.nd.empty_symmetric
    grads = [alloc(s) for s in shapes]
    if nval > 1:
        grads = [[mx.nd.empty(s, mx.gpu(local)) for _ in range(nval)] for s in shapes]
    weights = [alloc(s) for s in shapes]
    for g, s in zip(grads, shapes):
        for gg in (g if isinstance(g, list) else [g]):
            gg[:] = rng.uniform(-1, 1, s).astype(np.float32)
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(local)) for s in shapes])
    opt = os.environ.get("TUNE_OPT", "sgd")
    if opt == "sgd":
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.01, momentum=0.9, wd=1e-4))
    elif opt == "adam":
        kv.set_optimizer(mx.optimizer.Adam())
    configs = [(8192, 512, 0, 1)]
    steps = 40
    for chunk, threads, mb, bulk in configs:
        check_call(_LIB.MXKVB200SetTuning(ctypes.c_int64(chunk), threads, mb, bulk))
        for _ in range(5):
            kv.pushpull(keys, grads, out=weights)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            kv.pushpull(keys, grads, out=weights)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
        if rank == 0:
            if world == 1:
                bw = (S // 4) * (20 + 4 * nval) / (ms * 1e-3) / 1e9
                print("chunk %6d threads %3d max_blocks %4d bulk %d : %.4f ms  HBM %.0f GB/s" % (chunk, threads, mb, bulk, ms, bw), flush=True)
            else:
                bw = 2.0 * S * (world - 1) / world / (ms * 1e-3) / 1e9
                print("chunk %6d threads %3d max_blocks %4d bulk %d : %.4f ms  busbw %.0f GB/s" % (chunk, threads, mb, bulk, ms, bw), flush=True)
    mx.nd.waitall()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
