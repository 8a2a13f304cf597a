"""GPU parity of the dense reduce(+update) kernel against the CPU oracle, through the C ABI
(python ctypes front-end).  Bit-exact unless stated."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O

SIZES = [1, 7, 1000, 4099, 8192, 100003, (1 << 20) + 3]


def _rng(seed):
    return np.random.default_rng(1234 + seed)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def assert_bits_equal(a, b, msg=""):
    assert a.shape == b.shape, (a.shape, b.shape)
    ne = _bits(a) != _bits(b)
    assert not ne.any(), "%s: %d / %d elements differ, first at %s: %r vs %r" % (
        msg, ne.sum(), ne.size, np.argwhere(ne)[0], a[tuple(np.argwhere(ne)[0])], b[tuple(np.argwhere(ne)[0])])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 11])
@pytest.mark.parametrize("kvtype", ["device", "local"])
def test_sum_f32_same_device(n, kvtype):
    """push of n values resident on one GPU == CommDevice / CommCPU association order."""
    for E in SIZES:
        rng = _rng(E + n)
        vals = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        kv = mx.kv.create(kvtype)
        kv.init(0, mx.nd.zeros((E,), mx.gpu(0)))
        kv.push(0, [mx.nd.array(v, mx.gpu(0)) for v in vals])
        out = mx.nd.empty((E,), mx.gpu(0))
        kv.pull(0, out=out)
        if n == 1:
            want = vals[0]
        else:
            want = O.sum_device(vals) if kvtype == "device" else O.sum_cpu(vals)
        assert_bits_equal(out.asnumpy(), want, "n=%d E=%d %s" % (n, E, kvtype))


@pytest.mark.parametrize("dtype", [np.float64, np.int32, np.int64, np.uint8, np.int8, np.float16])
def test_sum_other_dtypes(dtype):
    for n in (2, 4, 5):
        for E in (3, 4099, 70001):
            rng = _rng(E * n)
            if np.dtype(dtype).kind == "f":
                vals = [rng.uniform(-1, 1, E).astype(dtype) for _ in range(n)]
            else:
                vals = [rng.integers(-100 if np.dtype(dtype).kind == "i" else 0, 100, E).astype(dtype)
                        for _ in range(n)]
            kv = mx.kv.create("device")
            kv.init("k", mx.nd.array(np.zeros(E, dtype), mx.gpu(0), dtype=dtype))
            kv.push("k", [mx.nd.array(v, mx.gpu(0), dtype=dtype) for v in vals])
            out = mx.nd.empty((E,), mx.gpu(0), dtype=dtype)
            kv.pull("k", out=out)
            assert_bits_equal(out.asnumpy(), O.sum_device(vals), "%s n=%d E=%d" % (dtype, n, E))


def test_sum_bf16_fp32_accumulate():
    """bf16 has no reference GPU path: fp32 accumulate in reduce order, one RNE."""
    for n in (2, 3, 8):
        for E in (5, 4099, 100001):
            rng = _rng(E + 7 * n)
            vals = [O.f32_to_bf16(rng.uniform(-1, 1, E).astype(np.float32)) for _ in range(n)]
            kv = mx.kv.create("device")
            kv.init(1, mx.nd.zeros((E,), mx.gpu(0), dtype="bfloat16"))
            kv.push(1, [mx.nd.array(v, mx.gpu(0), dtype="bfloat16") for v in vals])
            out = mx.nd.empty((E,), mx.gpu(0), dtype="bfloat16")
            kv.pull(1, out=out)
            want = O.sum_device(vals, bf16=True)
            assert_bits_equal(out.asnumpy(raw=True), want, "bf16 n=%d E=%d" % (n, E))
            # the stated tolerance of the north star (1e-2 relative) against the exact fp64 sum
            exact = sum(O.bf16_to_f32(v).astype(np.float64) for v in vals)
            got = O.bf16_to_f32(out.asnumpy(raw=True)).astype(np.float64)
            assert np.abs(got - exact).sum() <= 1e-2 * np.abs(exact).sum()


def _fused_case(opt_name, opt_kwargs, oracle_kwargs, n, E, steps, dtype=np.float32, cpu_vals=False):
    rng = _rng(E + n + steps)
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kv = mx.kv.create("device")
    kv.init(7, mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.create(opt_name, **opt_kwargs))
    okv = O.OracleKVStore("device")
    okv.init(7, w0.copy())
    okv.set_optimizer(O.OracleOptimizer(opt_name, **oracle_kwargs))
    out = mx.nd.empty((E,), mx.gpu(0))
    oout = np.empty(E, np.float32)
    for s in range(steps):
        grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        ctx = mx.cpu() if cpu_vals else mx.gpu(0)
        if s % 2 == 0:
            kv.push(7, [mx.nd.array(g, ctx) for g in grads])
            kv.pull(7, out=out)
        else:
            kv.pushpull(7, [mx.nd.array(g, ctx) for g in grads], out=out)
        okv.push(7, [g.copy() for g in grads])
        okv.pull(7, oout)
        assert_bits_equal(out.asnumpy(), oout, "%s step %d n=%d E=%d" % (opt_name, s, n, E))


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_fused_sgd(n):
    kw = dict(learning_rate=0.1, wd=1e-4, rescale_grad=1.0 / 32)
    for E in (9, 4099, 300007):
        _fused_case("sgd", kw, kw, n, E, 3)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_fused_sgd_momentum_clip(n):
    kw = dict(learning_rate=0.05, wd=1e-3, momentum=0.9, rescale_grad=0.5, clip_gradient=0.3)
    for E in (9, 4099, 300007):
        _fused_case("sgd", kw, kw, n, E, 4)


@pytest.mark.parametrize("n", [1, 3, 8])
def test_fused_adam(n):
    kw = dict(learning_rate=0.01, wd=1e-2, beta1=0.9, beta2=0.999, epsilon=1e-8, rescale_grad=0.25)
    for E in (9, 4099, 300007):
        _fused_case("adam", kw, kw, n, E, 4)


def test_fused_test_optimizer():
    kw = dict(learning_rate=0.3, wd=1e-2, rescale_grad=0.5)
    _fused_case("test", kw, kw, 4, 10007, 3)


def test_fused_host_values():
    """values living in (pageable) host memory are staged over PCIe and reduced on the GPU."""
    kw = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
    _fused_case("sgd", kw, kw, 4, 4099, 3, cpu_vals=True)


@pytest.mark.parametrize("pinned", [True, False])
def test_host_pipeline_segments(pinned, monkeypatch):
    """large host-resident values/outputs go through the segmented H2D / kernel / D2H pipeline;
    results must not depend on the segmentation."""
    monkeypatch.setenv("MXKV_B200_HOST_SEG_ELEMS", str(1 << 18))
    E, n = (1 << 20) + 12345, 3
    rng = _rng(77)
    hctx = mx.cpu_pinned() if pinned else mx.cpu()
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.05, momentum=0.9, wd=1e-4)
    kv = mx.kv.create("device")
    kv.init([0, 1], [mx.nd.array(w0, hctx), mx.nd.array(w0[:5000], hctx)])
    kv.set_optimizer(mx.optimizer.SGD(**kw))
    okv = O.OracleKVStore("device")
    okv.init([0, 1], [w0.copy(), w0[:5000].copy()])
    okv.set_optimizer(O.OracleOptimizer("sgd", **kw))
    outs = [mx.nd.empty((E,), hctx), mx.nd.empty((5000,), hctx)]
    for step in range(3):
        g0 = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        g1 = [rng.uniform(-1, 1, 5000).astype(np.float32) for _ in range(n)]
        kv.pushpull([0, 1], [[mx.nd.array(g, hctx) for g in g0], [mx.nd.array(g, hctx) for g in g1]], out=outs)
        okv.push([0, 1], [g0, g1])
        for k, o in enumerate(outs):
            o.wait_to_read()
            want = np.empty(o.shape, np.float32)
            okv.pull(k, want)
            assert_bits_equal(o.asnumpy(), want, "host pipeline step %d key %d" % (step, k))


@pytest.mark.parametrize("lp", ["bfloat16", np.float16])
def test_multi_precision_sgd_momentum(lp):
    """bf16/fp16 weights+grads, fp32 master and momentum (MP_SGDMomKernel)."""
    kind = 2 if lp == "bfloat16" else 1
    E, n = 50003, 4
    rng = _rng(99)
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    w0_lp = O.f32_to_bf16(w0) if kind == 2 else w0.astype(np.float16)
    w32 = O.bf16_to_f32(w0_lp) if kind == 2 else w0_lp.astype(np.float32)
    mom = np.zeros(E, np.float32)
    kv = mx.kv.create("device")
    kv.init(0, mx.nd.array(w0_lp, mx.gpu(0), dtype=lp))
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4, multi_precision=True))
    out = mx.nd.empty((E,), mx.gpu(0), dtype=lp)
    want_lp = np.zeros(E, np.uint16)
    for s in range(3):
        g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        g_lp = [O.f32_to_bf16(x) if kind == 2 else x.astype(np.float16) for x in g]
        kv.pushpull(0, [mx.nd.array(x, mx.gpu(0), dtype=lp) for x in g_lp], out=out)
        gsum = O.sum_device_lp_f32out(g_lp, kind)
        O.mp_sgd_mom_update(want_lp, kind, w32, mom, gsum, 0.1, 1e-4, 0.9)
        got = out.asnumpy(raw=True) if kind == 2 else out.asnumpy().view(np.uint16)
        assert_bits_equal(got, want_lp, "mp sgd %s step %d" % (lp, s))


def test_multi_key_one_launch():
    """a list of keys of mixed sizes goes through one launch and matches per-key results."""
    shapes = [(64,), (3, 5), (1000,), (257, 33), (2048, 16), (7,)]
    n = 4
    rng = _rng(5)
    kv = mx.kv.create("device")
    keys = list(range(len(shapes)))
    w0 = [rng.uniform(0, 1, s).astype(np.float32) for s in shapes]
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    kw = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
    kv.set_optimizer(mx.optimizer.SGD(**kw))
    okv = O.OracleKVStore("device")
    okv.init(keys, [w.copy() for w in w0])
    okv.set_optimizer(O.OracleOptimizer("sgd", **kw))
    outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    before = mx.kv.launch_count()
    for step in range(2):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(n)] for s in shapes]
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(0)) for g in gs] for gs in grads], out=outs)
        okv.push(keys, grads)
    assert mx.kv.launch_count() - before == 2, "one kernel launch per pushpull call expected"
    for k, o in zip(keys, outs):
        oo = np.empty(shapes[k], np.float32)
        okv.pull(k, oo)
        assert_bits_equal(o.asnumpy(), oo, "key %d" % k)


def test_unaligned_views_take_scalar_path():
    import torch
    E = 10001
    base = [torch.rand(E + 3, device="cuda") for _ in range(3)]
    views = [b[1:E + 1] for b in base]          # 4-byte aligned only
    kv = mx.kv.create("device")
    kv.init(0, mx.nd.zeros((E,), mx.gpu(0)))
    torch.cuda.synchronize()
    kv.push(0, [mx.nd.from_torch(v.contiguous()) if not v.is_contiguous() else mx.nd.from_torch(v) for v in views])
    out = mx.nd.empty((E,), mx.gpu(0))
    kv.pull(0, out=out)
    want = O.sum_device([v.cpu().numpy() for v in views])
    assert_bits_equal(out.asnumpy(), want)


def test_torch_stream_ordering():
    """gradients produced on torch's stream right before the call and weights consumed right
    after it: the event edges must order everything without host syncs."""
    import torch
    E = 1 << 22
    kv = mx.kv.create("device")
    w = torch.zeros(E, device="cuda")
    kv.init(0, mx.nd.from_torch(w))
    kv.set_optimizer(mx.optimizer.SGD(learning_rate=1.0))
    total = torch.zeros(E, device="cuda")
    for it in range(5):
        gs = [torch.full((E,), float(it + 1), device="cuda") * (k + 1) for k in range(3)]
        kv.pushpull(0, [mx.nd.from_torch(g) for g in gs], out=mx.nd.from_torch(w))
        total += sum(gs)           # consumes on torch's stream
        chk = (w + total).abs().max()
        assert chk.item() == 0.0, "iteration %d" % it


def test_save_load_optimizer_states(tmp_path):
    E = 5000
    rng = _rng(3)
    kw = dict(learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8)

    def run(kv, steps, seed):
        r = np.random.default_rng(seed)
        out = mx.nd.empty((E,), mx.gpu(0))
        for _ in range(steps):
            kv.pushpull(0, [mx.nd.array(r.uniform(-1, 1, E).astype(np.float32), mx.gpu(0)) for _ in range(2)], out=out)
        return out.asnumpy()

    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kv1 = mx.kv.create("device"); kv1.init(0, mx.nd.array(w0, mx.gpu(0))); kv1.set_optimizer(mx.optimizer.Adam(**kw))
    run(kv1, 2, 1)
    f = str(tmp_path / "states")
    kv1.save_optimizer_states(f, dump_optimizer=True)     # the optimizer carries the update counts
    w_mid = mx.nd.empty((E,), mx.gpu(0)); kv1.pull(0, out=w_mid)
    ref = run(kv1, 2, 2)
    kv2 = mx.kv.create("device"); kv2.init(0, w_mid); kv2.set_optimizer(mx.optimizer.Adam(**kw))
    kv2.load_optimizer_states(f)
    got = run(kv2, 2, 2)
    assert_bits_equal(got, ref)


# ---------------------------------------------------------------------------
# Forced kernel variants: the kernels bench.py times are the ones compared with the
# oracle here -- the shared-memory staged kernel (kv_dense_bulk_kernel) and the per-thread kernel
# (kv_dense_kernel) are each FORCED over BASELINE.json configs[1]'s sizes (2^10 ... 2^26 elements, one key
# per size and the whole key set in one call), n = 1 and 2 values per key on one GPU, no optimizer / SGD
# momentum / Adam, bit for bit; the variant launch counters prove which kernel ran.
# ---------------------------------------------------------------------------
import os

_SIM = bool(os.environ.get("MXKV_SIM"))
SWEEP_P = list(range(10, 19, 2)) if _SIM else list(range(10, 27, 2))      # the simulator runs on one CPU core
_OPTS = {
    "none": (None, {}),
    "sgd_mom": ("sgd", dict(learning_rate=0.01, momentum=0.9, wd=1e-4)),      # bench.py's optimizer
    "adam": ("adam", dict(learning_rate=0.001, wd=1e-3)),
}


class _forced(object):
    """force one dense kernel variant for the duration of a test; restores the automatic choice"""

    def __init__(self, variant):
        self.variant = variant

    def __enter__(self):
        mx.kv.set_tuning(bulk={"bulk": 2, "per_thread": 0}[self.variant])
        self.before = {v: mx.kv.launch_count(v) for v in ("per_thread", "bulk", "nvls")}
        return self

    def launched(self, v):
        return mx.kv.launch_count(v) - self.before[v]

    def __exit__(self, *a):
        mx.kv.set_tuning(bulk=1)


def _fill(rng, lo, hi, e):
    """random float32 data; large arrays repeat a random block of 1 000 003 elements (a period that no tile,
    chunk or shard boundary is a multiple of), because numpy draws only ~20 M values per second"""
    if e <= (1 << 20):
        return rng.uniform(lo, hi, e).astype(np.float32)
    return np.resize(rng.uniform(lo, hi, 1000003).astype(np.float32), e)


def _run_keyset(variant, optkey, n, sizes, steps=2, seed=0):
    optname, kw = _OPTS[optkey]
    rng = _rng(seed + n + len(sizes))
    keys = list(range(len(sizes)))
    w0 = [_fill(rng, 0, 1, e) for e in sizes]
    with _forced(variant) as f:
        kv = mx.kv.create("device")
        kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
        okv = O.OracleKVStore("device")
        okv.init(keys, [w.copy() for w in w0])
        if optname:
            kv.set_optimizer(mx.optimizer.create(optname, **kw))
            okv.set_optimizer(O.OracleOptimizer(optname, **kw))
        outs = [mx.nd.empty((e,), mx.gpu(0)) for e in sizes]
        gdev = [[mx.nd.empty((e,), mx.gpu(0)) for _ in range(n)] for e in sizes]
        big = sum(sizes) > (1 << 24)
        for s in range(steps):
            if s == 0 or not big:      # large key sets push the same gradients again (the state has moved on)
                grads = [[_fill(rng, -1, 1, e) for _ in range(n)] for e in sizes]
                for gd, gs in zip(gdev, grads):
                    for d, g in zip(gd, gs):
                        d[:] = g
            kv.pushpull(keys, gdev, out=outs)
            okv.push(keys, grads)
            for k, e in enumerate(sizes):
                want = np.empty(e, np.float32)
                okv.pull(k, want)
                assert_bits_equal(outs[k].asnumpy(), want, "%s %s n=%d E=%d step %d" % (variant, optkey, n, e, s))
        mx.nd.waitall()
        other = "per_thread" if variant == "bulk" else "bulk"
        assert f.launched(variant) >= steps, "the %s kernel did not run (%d launches)" % (variant, f.launched(variant))
        assert f.launched(other) == 0, "%d launches of the %s kernel under a forced %s" % (f.launched(other), other, variant)


@pytest.mark.parametrize("optkey", ["none", "sgd_mom", "adam"])
@pytest.mark.parametrize("n", [1, 2])
@pytest.mark.parametrize("variant", ["bulk", "per_thread"])
def test_forced_variant_every_sweep_size(variant, n, optkey):
    for p in [q for q in SWEEP_P if q <= 22]:      # 2^24 and 2^26: the whole-key-set test below
        _run_keyset(variant, optkey, n, [1 << p], steps=2, seed=p)


@pytest.mark.parametrize("optkey", ["none", "sgd_mom", "adam"])
@pytest.mark.parametrize("n", [1, 2])
@pytest.mark.parametrize("variant", ["bulk", "per_thread"])
def test_forced_variant_whole_sweep_key_set_in_one_call(variant, n, optkey):
    """the bench workload itself: 9 keys 4 KB ... 256 MB in ONE pushpull = one launch"""
    _run_keyset(variant, optkey, n, [1 << p for p in SWEEP_P], steps=2, seed=77)


@pytest.mark.parametrize("variant", ["bulk", "per_thread"])
def test_forced_variant_ragged_tiles(variant):
    """sizes that are multiples of 4 but not of the 2048-element tile / 8192-element chunk: partial last tiles,
    tiles of 4 elements, a key smaller than one tile next to a large one"""
    _run_keyset(variant, "sgd_mom", 2, [4, 2052, 8196, 3 * 2048 + 12, (1 << 20) + 4, 12], steps=3, seed=5)
    _run_keyset(variant, "adam", 1, [2048, 2044, 4096 + 8, 100004], steps=2, seed=6)


# ---------------------------------------------------------------------------
# cached launch plans: the third identical pushpull in a row replays the recorded work lists
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("optkey", ["none", "sgd_mom", "adam"])
def test_cached_launch_plan_replays_are_bit_exact(optkey):
    optname, kw = _OPTS[optkey]
    sizes = [8, 1000, 4096 + 4, 70001, 1 << 18]
    rng = _rng(31)
    keys = list(range(len(sizes)))
    w0 = [rng.uniform(0, 1, e).astype(np.float32) for e in sizes]
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    okv = O.OracleKVStore("device")
    okv.init(keys, [w.copy() for w in w0])
    if optname:
        kv.set_optimizer(mx.optimizer.create(optname, **kw))
        okv.set_optimizer(O.OracleOptimizer(optname, **kw))
    n = 2
    gdev = [[mx.nd.empty((e,), mx.gpu(0)) for _ in range(n)] for e in sizes]
    outs = [mx.nd.empty((e,), mx.gpu(0)) for e in sizes]

    def step(check=True):
        grads = [[rng.uniform(-1, 1, e).astype(np.float32) for _ in range(n)] for e in sizes]
        for gd, gs in zip(gdev, grads):
            for d, g in zip(gd, gs):
                d[:] = g
        kv.pushpull(keys, gdev, out=outs)
        okv.push(keys, grads)
        for k, e in enumerate(sizes):
            want = np.empty(e, np.float32)
            okv.pull(k, want)
            assert_bits_equal(outs[k].asnumpy(), want, "%s key %d" % (optkey, k))

    h0 = kv.plan_hits()
    for _ in range(6):                  # calls 1, 2 build (2 records the plan); 3 ... 6 replay it
        step()
    assert kv.plan_hits() - h0 == 4, kv.plan_hits() - h0
    # anything else that touches a key expires the plan: a pull in between, then two more calls to re-record
    tmp = mx.nd.empty((sizes[1],), mx.gpu(0))
    kv.pull(1, out=tmp)
    h1 = kv.plan_hits()
    step(); step()
    assert kv.plan_hits() == h1
    step()
    assert kv.plan_hits() == h1 + 1
    # a learning-rate change between replays takes effect (the scalars are patched at every call)
    if optname:
        h2 = kv.plan_hits()
        kv._optimizer.set_learning_rate(kv._optimizer.learning_rate * 0.5)
        okv.optimizer.lr = okv.optimizer.lr * 0.5
        step()
        assert kv.plan_hits() == h2 + 1
