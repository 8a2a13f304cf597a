#!/usr/bin/env python
"""bench.py -- KVStore push+pull throughput on B200 (BASELINE.json metric).

A *step* is one pass of the hot path over one batch of synthetic gradients: `pushpull` of the
whole key set (sum over the ranks' gradients, fused SGD-momentum update of the stored fp32
weights, new weights delivered to every rank).  Workload (config.workload):

  sweep     BASELINE.json configs[1]: one fp32 key per size 2^10, 2^12, ..., 2^26 elements
            (4 KB ... 256 MB; 89.5 M elements, 358 MB per rank) -- the default
  resnet50  configs[2]'s key set: the 193 gradient-carrying arrays of Gluon ResNet-50-v1 (25.6 M elements)
  bert      configs[3]'s key set: BERT-base (synthetic shape list), Adam

One process per GPU (torchrun for N > 1); every rank contributes its own gradient for every key
(weak scaling: per-GPU work is fixed).  `value` = N * 2 * S / t  [GB/s]: bytes pushed plus bytes
pulled by all ranks per second, S = key-set bytes.  Inputs are larger than L2 (358 MB of
gradients + 358 MB of weights + state per step vs 126 MB of L2), so no flush is needed.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
def keyset(workload):
    if workload == "sweep":
        return [(1 << p,) for p in range(10, 27, 2)]
    if workload == "resnet50":
        # python/mxnet/gluon/model_zoo/vision/resnet.py:345-422 (BottleneckV1, layers [3,4,6,3],
        # channels [64,256,512,1024,2048]); conv weights, the biases Gluon leaves on the 1x1
        # convs, BN gamma/beta, FC weight/bias
        shapes = [(64, 3, 7, 7), (64,), (64,)]
        inp = 64
        for stage, (blocks, ch) in enumerate(zip([3, 4, 6, 3], [256, 512, 1024, 2048])):
            mid = ch // 4
            for b in range(blocks):
                shapes += [(mid, inp, 1, 1), (mid,), (mid,), (mid,)]          # conv1x1 + bias + BN
                shapes += [(mid, mid, 3, 3), (mid,), (mid,)]                  # conv3x3 + BN
                shapes += [(ch, mid, 1, 1), (ch,), (ch,), (ch,)]              # conv1x1 + bias + BN
                if b == 0:
                    shapes += [(ch, inp, 1, 1), (ch,), (ch,)]                 # downsample conv + BN
                inp = ch
        shapes += [(1000, 2048), (1000,)]
        return shapes
    if workload == "bert":
        H, FF = 768, 3072
        shapes = [(30522, H), (512, H), (2, H), (H,), (H,)]
        for _ in range(12):
            shapes += [(H, H), (H,)] * 3 + [(H, H), (H,), (H,), (H,), (FF, H), (FF,), (H, FF), (H,), (H,), (H,)]
        shapes += [(H, H), (H,)]
        return shapes
    raise ValueError(workload)


def nelem(shape):
    n = 1
    for d in shape:
        n *= d
    return n


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU KVStore ('local': CommCPU reduce +
# CPU updater + ParallelCopy broadcast) restated in oracle/kv_oracle.c, all host cores
# ---------------------------------------------------------------------------
def cpu_kvstore_step_factory(shapes, n_values, threads, optimizer):
    from oracle import oracle as O
    rng = np.random.default_rng(1234)
    sizes = [nelem(s) for s in shapes]
    grads = [[rng.uniform(-1, 1, e).astype(np.float32) for _ in range(n_values)] for e in sizes]
    stage = [[np.empty(e, np.float32) for _ in range(n_values)] for e in sizes]    # pinned merge/copy bufs
    weights = [rng.uniform(0, 1, e).astype(np.float32) for e in sizes]
    mom = [np.zeros(e, np.float32) for e in sizes]
    mean = [np.zeros(e, np.float32) for e in sizes] if optimizer in ("adam", "lamb", "lans") else None
    outs = [[np.empty(e, np.float32) for _ in range(n_values)] for e in sizes]
    lib = O.lib()
    import ctypes
    state = {"t": 0}
    lib.kvo_set_threads(ctypes.c_int(threads))     # the optimizer loops use the OpenMP default team

    def step():
        state["t"] += 1
        for k, e in enumerate(sizes):
            if n_values == 1:
                merged = grads[k][0]                                  # comm.h:128-131
            else:
                for j in range(n_values):                             # CopyFromTo(src[j], &buf) comm.h:147,162
                    lib.kvo_parallel_copy_f32(stage[k][j].ctypes.data_as(O.c_f32p),
                                              grads[k][j].ctypes.data_as(O.c_f32p), ctypes.c_int64(e),
                                              ctypes.c_int(threads))
                O.sum_cpu_inplace(stage[k], nthreads=threads)         # ReduceSumCPU comm.h:359-411
                merged = stage[k][0]
            if optimizer == "adam":
                O.adam_update(weights[k], merged, mean[k], mom[k], O.adam_lr(0.001, 0.9, 0.999, state["t"]))
            elif optimizer == "sgd":
                O.sgd_mom_update(weights[k], merged, mom[k], 0.01, 1e-4, 0.9)
            elif optimizer == "lamb":        # multi_lamb_update on the CPU: sequential norms + two passes
                O.lamb_update(weights[k], merged, mean[k], mom[k], 0.001, 0.01, state["t"])
            elif optimizer == "lans":
                O.lans_update(weights[k], merged, mean[k], mom[k], 0.001, 0.01, state["t"])
            elif optimizer == "lars":
                O.sgd_mom_update(weights[k], merged, mom[k], O.lars_lr(0.1, weights[k], merged, 1e-4), 1e-4, 0.9)
            else:
                weights[k][...] = merged
            for j in range(n_values):                                 # Broadcast: CopyFromTo(local, out_j)
                lib.kvo_parallel_copy_f32(outs[k][j].ctypes.data_as(O.c_f32p),
                                          weights[k].ctypes.data_as(O.c_f32p), ctypes.c_int64(e),
                                          ctypes.c_int(threads))
    return step


def ncu_dram_bytes(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first kernel in an `ncu --page raw` text dump."""
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    got = {}
    try:
        with open(path) as f:
            for line in f:
                parts = line.split()
                if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") \
                        and parts[0] not in got and parts[-1] in units:
                    got[parts[0]] = float(parts[-2].replace(",", "")) * units[parts[-1]]
    except OSError:
        return None
    return sum(got.values()) if len(got) == 2 else None


def host_cpus():
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:            # cgroup v2
            q, per = f.read().split()
            if q != "max":
                n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:   # cgroup v1
                q = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                per = int(f.read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def run_cpu_reference(args, shapes, as_baseline=False):
    """Times the reference's CPU KVStore algorithm (oracle port, all host threads).  A step is a
    BOUNDED SAMPLE of the workload: every key is cut to the same leading fraction so that the whole
    --steps/--warmup run stays within ~2 minutes (GB/s is intensive, so the sample is comparable)."""
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")     # oversubscribed spinning would only hurt the CPU arm
    from oracle import oracle as O
    avail = max(1, min(O.lib().kvo_max_threads(), host_cpus()))
    n_values = max(1, args.gpus)
    warm = 1 if as_baseline else args.warmup
    steps = 3 if as_baseline else args.steps
    budget_s = 20.0 if as_baseline else 100.0
    sizes = [nelem(s) for s in shapes]
    probe_frac = min(1.0, 64e6 / max(1, sum(sizes)) / n_values * 4)     # probe on <= ~64 M elements of traffic
    probe_shapes = [(max(256, int(e * probe_frac)),) for e in sizes]
    # give the CPU arm its best thread count (all usable cores, or fewer when memory-bound)
    threads, t_probe = avail, None
    for cand in sorted({avail, max(1, avail // 2), max(1, avail // 4)}, reverse=True):
        probe = cpu_kvstore_step_factory(probe_shapes, n_values, cand, args.optimizer)
        probe()
        t0 = time.perf_counter(); probe(); tp = time.perf_counter() - t0
        if t_probe is None or tp < t_probe:
            threads, t_probe = cand, tp
    t_full = t_probe / probe_frac
    frac = min(1.0, budget_s / (steps + warm) / t_full)
    sample_shapes = [(max(256, int(e * frac)),) for e in sizes]
    S = 4 * sum(nelem(s) for s in sample_shapes)
    step = cpu_kvstore_step_factory(sample_shapes, n_values, threads, args.optimizer)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = n_values * 2 * S / dt / 1e9
    sample = ("%d step(s), each over the leading %.1f%% of every key of the %s key set (%.1f MB x %d value(s) per key), "
              "%d OpenMP threads" % (steps, 100 * frac, args.workload, S / 1e6, n_values, threads))
    base = {"value": value, "unit": "GB/s", "cores": threads,
            "kind": "port", "sample": sample}
    return value, dt, base


# ---------------------------------------------------------------------------
# samples/sec: BASELINE.json configs[2] -- ResNet-50 bf16, synthetic 3x224x224, SGD-momentum
# (multi-precision) through Trainer(kvstore='device'); torch does forward/backward, the engine does
# the gradient exchange + fused update.  Secondary workload: one JSON line with img/s.
# ---------------------------------------------------------------------------
def train_resnet50(args):
    import torch
    import torchvision
    import mxnet_b200 as mx
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        mx.dist.init_process_group(device=local)
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).cuda().to(memory_format=torch.channels_last).to(torch.bfloat16)
    params = [p for p in model.parameters() if p.requires_grad]
    trainer = mx.Trainer(params, "sgd", {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4, "multi_precision": True},
                         kvstore="device", symmetric=True)
    B = args.batch
    x = torch.randn(B, 3, 224, 224, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (B,), device="cuda")
    lossf = torch.nn.CrossEntropyLoss()

    def step():
        out = model(x)
        loss = lossf(out.float(), y)
        loss.backward()
        trainer.step(B * world)
        return loss

    steps, warm = min(args.steps, 50), max(3, args.warmup)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ec = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = mx.kv.launch_count()
    e0.record()
    for i in range(steps):
        out = model(x)
        loss = lossf(out.float(), y)
        loss.backward()
        ec[i][0].record()
        trainer.step(B * world)
        ec[i][1].record()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    comm_ms = sum(a.elapsed_time(b) for a, b in ec) / steps
    t = torch.tensor([ms, comm_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, comm_ms = t[0].item(), t[1].item()
    nparam = sum(p.numel() for p in params)
    if rank == 0:
        print(json.dumps({"metric": "samples/sec (ResNet-50 bf16, synthetic 3x224x224, Trainer(kvstore='device'))",
                          "value": world * B / (ms * 1e-3), "unit": "img/s", "n_gpus": world, "steps": steps,
                          "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": "resnet50-train: torchvision resnet50 (random init), batch %d per GPU, "
                                                 "SGD momentum 0.9 wd 1e-4 multi-precision, %d keys / %.1f M params"
                                                 % (B, len(params), nparam / 1e6),
                                     "exchange_ms_per_step": comm_ms, "keys": len(params)},
                          "gpu_launches": mx.kv.launch_count() - launches0, "loss": float(loss)}))
    mx.nd.waitall()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------
# BASELINE.json configs[4]: one row_sparse key (1 M x 256 fp32), 10 k distinct rows per GPU;
# step = row_sparse push (union + gather-sum + lazy SGD-momentum on the touched rows) followed by
# row_sparse_pull of the same ids.  Secondary workload: one JSON line with rows/s and GB/s.
# ---------------------------------------------------------------------------
def bench_rsp(args):
    import torch
    import mxnet_b200 as mx
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        mx.dist.init_process_group(device=local)
    ctx = mx.gpu(local)
    R, L, nnz = 1_000_000, 256, 10_000
    rng = np.random.default_rng(1234 + rank)
    idx = np.sort(rng.choice(R, nnz, replace=False)).astype(np.int64)
    val = rng.uniform(-1, 1, (nnz, L)).astype(np.float32)
    grad = mx.nd.row_sparse_array((val, idx), shape=(R, L), ctx=ctx)
    ids = mx.nd.array(idx, ctx, dtype=np.int64)
    out = mx.nd.empty((R, L), ctx, stype="row_sparse", capacity=nnz)
    kv = mx.kv.create("device")
    kv.init("emb", mx.nd.row_sparse_array((np.zeros((1, L), np.float32), np.zeros(1, np.int64)), shape=(R, L), ctx=ctx))
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.01, momentum=0.9, wd=0.0, lazy_update=True))

    def step():
        kv.push("emb", grad)
        kv.row_sparse_pull("emb", out=out, row_ids=ids)

    steps, warm = min(args.steps, 100), max(3, args.warmup)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = mx.kv.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = t.item()
    union = int(out.indices.shape[0])
    row_bytes = L * 4 + 8
    alg_per_gpu = world * nnz * row_bytes + nnz * row_bytes    # every GPU reads all sources' rows, then pulls its rows
    if rank == 0:
        print(json.dumps({"metric": "row_sparse push + row_sparse_pull rows/s", "value": world * nnz / (ms * 1e-3),
                          "unit": "rows/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": "rsp: 1 key 1000000x256 fp32 row_sparse, %d rows per GPU, lazy SGD-momentum, "
                                                 "push + row_sparse_pull" % nnz,
                                     "union_rows": union,
                                     "algorithmic_gbs_per_gpu": alg_per_gpu / (ms * 1e-3) / 1e9},
                          "gpu_launches": mx.kv.launch_count() - launches0}))
    mx.nd.waitall()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "resnet50", "bert", "resnet50-train", "rsp"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch of resnet50-train")
    ap.add_argument("--optimizer", default=None, choices=[None, "sgd", "adam", "lamb", "lans", "lars", "none"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nvls", action="store_true", help="keep gradients/weights out of multicast memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--local-world", type=int, default=0,
                    help="split the box into 'nodes' of this many GPUs and use kv.create('dist_device_sync'): NVLink "
                         "peer memory inside a node, NCCL between nodes (not the driver's configuration)")
    args = ap.parse_args()
    if args.workload == "resnet50-train":
        return train_resnet50(args)
    if args.workload == "rsp":
        return bench_rsp(args)
    if args.optimizer is None:
        args.optimizer = "adam" if args.workload == "bert" else "sgd"
    shapes = keyset(args.workload)
    S = 4 * sum(nelem(s) for s in shapes)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    config = {"workload": "%s: pushpull of %d fp32 keys (%.1f MB per rank), fused %s update" % (
        args.workload, len(shapes), S / 1e6, args.optimizer),
        "keys": len(shapes), "bytes_per_rank": S, "optimizer": args.optimizer,
        "parallelism": "dp%d (one process per GPU, gradient exchange over NVLink peer memory)" % args.gpus,
        "l2": "working set (%.0f MB grads + weights + state per rank) exceeds the 126 MB L2; no flush" % (3 * S / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        value, dt, base = run_cpu_reference(args, shapes)
        line = {"impl": "reference", "metric": "kvstore push+pull GB/s", "value": value, "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import mxnet_b200 as mx
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        hier = 0 < args.local_world < world
        mx.dist.init_process_group(device=local, local_world=args.local_world if hier else world)
        dev = local
    else:
        assert args.gpus == 1, "launch with torchrun for --gpus > 1"
        dev = 0
        hier = False
        torch.cuda.set_device(0)
    ctx = mx.gpu(dev)

    rng = np.random.default_rng(1234 + rank)
    keys = list(range(len(shapes)))
    # gradients / weights live in the peer-mapped arena (zero-copy over NVLink); at N=1 this is
    # plain device memory
    exchange = "nvlink-p2p"
    alloc = mx.nd.empty_symmetric
    if hier:
        exchange = "hierarchical: nvlink-p2p in nodes of %d, nccl between %d nodes" % (
            args.local_world, world // args.local_world)
    if world > 1 and not args.no_nvls and not hier:
        try:
            probe = mx.nd.empty_multicast((1024,))
            if mx.nd.has_multicast(probe):
                alloc = mx.nd.empty_multicast          # NVSwitch multicast-capable arrays
                # the engine switches to the multimem kernel above 4 ranks (MXKVB200SetNvls)
                exchange = "nvls-multicast" if world > 4 else "nvlink-p2p (multicast-capable arrays)"
        except Exception as e:                         # noqa: BLE001 -- torch symmetric memory unavailable
            sys.stderr.write("multicast allocation unavailable (%r): peer-load kernels\n" % (e,))
    grads = [alloc(s) for s in shapes]
    weights = [alloc(s) for s in shapes]
    for g, s in zip(grads, shapes):
        g[:] = rng.uniform(-1, 1, s).astype(np.float32)
    config["exchange"] = exchange
    w0 = np.random.default_rng(99)
    kv = mx.kv.create("dist_device_sync" if hier else "device")
    kv.init(keys, [mx.nd.array(w0.uniform(0, 1, s).astype(np.float32), ctx) for s in shapes])
    if args.optimizer == "sgd":
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.01, momentum=0.9, wd=1e-4))
        bytes_per_elem_n1 = 6 * 4      # read g, w, mom; write w(store), mom, w(out)
    elif args.optimizer == "adam":
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=0.001))
        bytes_per_elem_n1 = 8 * 4      # read g, w, m, v; write w, m, v, out
    elif args.optimizer == "lamb":
        kv.set_optimizer(mx.optimizer.LAMB(learning_rate=0.001, wd=0.01))
        # first: read g, w, m, v; write m, v, ghat.  apply: read w, ghat; write w(store), out
        bytes_per_elem_n1 = 11 * 4
    elif args.optimizer == "lans":
        kv.set_optimizer(mx.optimizer.LANS(learning_rate=0.001, wd=0.01))
        # first: read g, w; write g.  mid: read g, w, m, v; write m, v, temp_m, temp_g.
        # apply: read w, temp_m, temp_g; write w(store), out
        bytes_per_elem_n1 = 16 * 4
    elif args.optimizer == "lars":
        kv.set_optimizer(mx.optimizer.LARS(learning_rate=0.1, momentum=0.9, wd=1e-4))
        # first: read g, w; write g.  apply: read w, g, mom; write w(store), mom, out
        bytes_per_elem_n1 = 9 * 4
    else:
        bytes_per_elem_n1 = 3 * 4      # read g; write store, out
    layerwise = args.optimizer in ("lamb", "lans", "lars")

    def step():
        kv.pushpull(keys, grads, out=weights)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    import ctypes
    from mxnet_b200.base import _LIB, check_call
    sp = ctypes.c_void_p()
    check_call(_LIB.MXKVB200GetEngineStream(dev, ctypes.byref(sp)))
    engine_stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", dev))

    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    barrier()

    # ---- timed region: device time (inputs resident in HBM) ----------------------------------
    launches0 = mx.kv.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize(); barrier()
    ev0.record()
    for i in range(args.steps):
        kern[i][0].record(engine_stream)
        step()
        kern[i][1].record(engine_stream)
    ev1.record()
    torch.cuda.synchronize(); barrier()
    launches = mx.kv.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in kern)
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / args.steps
    kern_ms = t[1].item()
    value = world * 2 * S / (ms_step * 1e-3) / 1e9

    # ---- roofline for the dominant (only) kernel ------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    nel = S // 4
    if world == 1:
        alg = nel * bytes_per_elem_n1
        peak = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "hbm", "achieved": alg / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "traffic": None, "kernel": "kv_norm_first/(mid)/apply_kernel sequence" if layerwise
                else "kv_dense_bulk_kernel", "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg}
    else:
        nvls_active = exchange == "nvls-multicast" and world > 4
        if nvls_active:
            # in-switch reduction + replication: per GPU per direction S (own data out / all shards in)
            # plus S/n (its reduced shard in / its updated shard out)
            alg = S * (1.0 + 1.0 / world)
            kname = "kv_dense_nvls_kernel (multimem.ld_reduce + fused update + multimem.st)"
        else:
            alg = 2.0 * S * (world - 1) / world      # per GPU per direction (tools/bandwidth/measure.py:138)
            kname = "kv_dense_bulk_kernel (peer cp.async.bulk loads + fused update + peer stores)"
        roof = {"bound": "nvlink", "achieved": alg / (kern_ms * 1e-3) / 1e9, "peak": 770.0, "unit": "GB/s",
                "peak_source": "B200_PROFILING.md measured peer copy 770 GB/s/dir (900 nominal)",
                "traffic": None, "kernel": kname, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg,
                "note": "bytes that must cross this GPU's NVLink per direction; busbw_gbs_per_gpu is the "
                        "NCCL-comparable 2S(n-1)/n / t"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    if world == 1 and args.workload == "sweep" and args.optimizer == "sgd":
        # DRAM bytes of the same kernel on the same workload from this round's `ncu --set full` capture
        # (bench.py cannot run under ncu itself: a number printed under a profiler is never a bench value)
        src = os.path.join(ROOT, "profiles", "r01_final_n1_kv_dense_bulk_kernel_ncu_full.txt")
        roof["traffic"], roof["traffic_source"] = ncu_dram_bytes(src), os.path.relpath(src, ROOT)

    # ---- e2e: host gradients in, host weights out, through the same public API ---------------------
    e2e = None
    if not args.no_e2e:
        hg = [mx.nd.empty(s, mx.cpu_pinned()) for s in shapes]
        hw = [mx.nd.empty(s, mx.cpu_pinned()) for s in shapes]
        for g, s in zip(hg, shapes):
            g[:] = rng.uniform(-1, 1, s).astype(np.float32)
        for _ in range(2):
            kv.pushpull(keys, hg, out=hw)
        mx.nd.waitall(); torch.cuda.synchronize(); barrier()
        n_e2e = max(3, args.steps // 4)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            kv.pushpull(keys, hg, out=hw)
        mx.nd.waitall(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item() / n_e2e
        e2e = {"value": world * 2 * S / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": S,
               "d2h_bytes_per_step": S, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "how": "pinned host gradients -> kv.pushpull -> pinned host weights, wall clock incl. both copies"}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            mx.nd.waitall()
        return

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        _, _, cpu_base = run_cpu_reference(args, shapes, as_baseline=True)

    line = {"metric": "kvstore push+pull GB/s", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": roof, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
            "busbw_gbs_per_gpu": (2.0 * S * (world - 1) / world) / (ms_step * 1e-3) / 1e9 if world > 1 else 0.0}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


if __name__ == "__main__":
    main()
