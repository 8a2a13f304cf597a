"""1-bit / 2-bit gradient compression with error feedback (tests/nightly/test_kvstore.py:121-295)
against the oracle codec (itself pinned to the reference's bit-level simulator, tests/golden)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("kind,thr", [("2bit", 0.5), ("2bit", 0.2), ("1bit", 0.0), ("1bit", 0.3)])
@pytest.mark.parametrize("E,n", [(16, 1), (100, 4), (4099, 3), (65536, 2)])
def test_compressed_push_matches_oracle(kind, thr, E, n):
    rng = np.random.default_rng(E + n)
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": kind, "threshold": thr})
    kv.init(0, mx.nd.zeros((E,), mx.gpu(0)))
    residual = [np.zeros(E, np.float32) for _ in range(n)]
    out = mx.nd.empty((E,), mx.gpu(0))
    for step in range(4):
        grads = [rng.uniform(-1.2, 1.2, E).astype(np.float32) for _ in range(n)]
        kv.push(0, [mx.nd.array(g, mx.gpu(0)) for g in grads])
        kv.pull(0, out=out)
        deq = []
        for g, r in zip(grads, residual):
            if kind == "2bit":
                deq.append(O.dequantize_2bit(O.quantize_2bit(g, r, thr), E, thr))
            else:
                deq.append(O.dequantize_1bit(O.quantize_1bit(g, r, thr), E, thr))
        want = O.sum_device(deq) if n > 1 else deq[0]
        assert _bits_equal(out.asnumpy(), want), (kind, step)


def test_nightly_2bit_known_answers():
    """tests/nightly/test_kvstore.py:150-200: zeros push -> 0; ones*threshold... with threshold 0.5:
    pushing 0.3 twice stays 0 then fires once the residual reaches the threshold."""
    thr = 0.5
    shape = (3, 5)
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": "2bit", "threshold": thr})
    kv.init(1, mx.nd.zeros(shape, mx.gpu(0)))
    out = mx.nd.empty(shape, mx.gpu(0))
    kv.push(1, [mx.nd.zeros(shape, mx.gpu(0)) for _ in range(2)])
    kv.pull(1, out=out)
    assert np.all(out.asnumpy() == 0)
    kv.push(1, [mx.nd.ones(shape, mx.gpu(0)) * 0.3 for _ in range(2)])
    kv.pull(1, out=out)
    assert np.all(out.asnumpy() == 0)              # residual 0.3 < threshold
    kv.push(1, [mx.nd.ones(shape, mx.gpu(0)) * 0.3 for _ in range(2)])
    kv.pull(1, out=out)
    assert np.all(out.asnumpy() == 2 * thr)        # residual 0.6 >= threshold on both "devices"
    kv.push(1, [mx.nd.ones(shape, mx.gpu(0)) * -0.7 for _ in range(2)])
    kv.pull(1, out=out)
    assert np.all(out.asnumpy() == -2 * thr)       # residual 0.1 - 0.7 = -0.6 <= -threshold


def test_compressed_fused_sgd():
    E, n, thr = 10000, 4, 0.5
    rng = np.random.default_rng(0)
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": "2bit", "threshold": thr})
    kv.init(0, mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.SGD(**kw))
    okv = O.OracleKVStore("device")
    okv.init(0, w0.copy())
    okv.set_optimizer(O.OracleOptimizer("sgd", **kw))
    residual = [np.zeros(E, np.float32) for _ in range(n)]
    out = mx.nd.empty((E,), mx.gpu(0))
    for step in range(3):
        grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        kv.pushpull(0, [mx.nd.array(g, mx.gpu(0)) for g in grads], out=out)
        deq = [O.dequantize_2bit(O.quantize_2bit(g, r, thr), E, thr) for g, r in zip(grads, residual)]
        okv.push(0, deq)
        want = np.empty(E, np.float32)
        okv.pull(0, want)
        assert _bits_equal(out.asnumpy(), want), step


@pytest.mark.multigpu
def test_compressed_multi_gpu():
    devs = list(range(min(mx.num_gpus(), 4)))
    E, thr = 50000, 0.5
    rng = np.random.default_rng(1)
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": "2bit", "threshold": thr})
    kv.init(0, mx.nd.zeros((E,), mx.gpu(0)))
    residual = [np.zeros(E, np.float32) for _ in devs]
    for step in range(3):
        grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
        outs = [mx.nd.empty((E,), mx.gpu(d)) for d in devs]
        kv.pushpull(0, [mx.nd.array(g, mx.gpu(d)) for g, d in zip(grads, devs)], out=outs)
        deq = [O.dequantize_2bit(O.quantize_2bit(g, r, thr), E, thr) for g, r in zip(grads, residual)]
        want = O.sum_device(deq)
        for o in outs:
            assert _bits_equal(o.asnumpy(), want), step
