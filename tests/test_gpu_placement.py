"""Placement edge cases of the single-process multi-GPU shape: which GPUs take part, where the stored value
and the optimizer state live, what happens when the set of devices changes between calls, outputs on GPUs
that did not push, host and device values mixed, more destinations than one launch can address, in-place
forms.  Everything is compared bit-for-bit with the oracle.  On hardware these need several GPUs (skipped
otherwise); the CPU suite runs them on 8 simulated GPUs (tests/test_sim_host_logic.py)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

import mxnet_b200 as mx
from oracle import oracle as O


def _need(n):
    if mx.num_gpus() < n:
        pytest.skip("needs %d GPUs" % n)


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _rng(seed):
    return np.random.default_rng(777 + seed)


@pytest.mark.parametrize("E", [1000, 300007])          # one-shot and two-shot sizes
def test_subset_of_gpus_and_outputs_elsewhere(E):
    """values pushed from GPUs 1 and 3 only; outputs on 0, 2 (not participants) and on the host"""
    _need(4)
    rng = _rng(E)
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
    kv = mx.kv.create("device")
    kv.init(0, mx.nd.array(w0, mx.gpu(2)))
    kv.set_optimizer(mx.optimizer.SGD(**kw))
    okv = O.OracleKVStore("device"); okv.init(0, w0.copy()); okv.set_optimizer(O.OracleOptimizer("sgd", **kw))
    outs = [mx.nd.empty((E,), mx.gpu(0)), mx.nd.empty((E,), mx.gpu(2)), mx.nd.empty((E,), mx.cpu()),
            mx.nd.empty((E,), mx.gpu(3))]
    want = np.empty(E, np.float32)
    for step in range(3):
        g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(2)]
        kv.pushpull(0, [mx.nd.array(g[0], mx.gpu(1)), mx.nd.array(g[1], mx.gpu(3))], out=outs)
        okv.push(0, g); okv.pull(0, want)
        for o in outs:
            assert _bits_equal(o.asnumpy(), want), (step, o.context)


def test_device_set_changes_between_calls():
    """2 GPUs, then 4, then 1, then 3 different ones: the stored value and the (sharded or replicated) Adam
    state follow -- the reference keeps them on one root and does not care; here they are re-gathered."""
    _need(4)
    E = 300007
    rng = _rng(1)
    w0 = rng.uniform(0, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.01, wd=1e-3)
    kv = mx.kv.create("device")
    kv.init("w", mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.Adam(**kw))
    okv = O.OracleKVStore("device"); okv.init("w", w0.copy()); okv.set_optimizer(O.OracleOptimizer("adam", **kw))
    want = np.empty(E, np.float32)
    for devs in ([0, 1], [0, 1, 2, 3], [2], [3, 1, 0], [0, 1]):
        g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
        outs = [mx.nd.empty((E,), mx.gpu(d)) for d in devs]
        kv.pushpull("w", [mx.nd.array(x, mx.gpu(d)) for x, d in zip(g, devs)], out=outs)
        okv.push("w", g); okv.pull("w", want)
        for o in outs:
            assert _bits_equal(o.asnumpy(), want), devs
    pulled = mx.nd.empty((E,), mx.gpu(3))
    kv.pull("w", out=pulled)
    assert _bits_equal(pulled.asnumpy(), want)


def test_host_and_device_values_mixed():
    """one value in host memory, the others on two GPUs: staged onto the root, summed in push order"""
    _need(2)
    E = 50003
    rng = _rng(2)
    kv = mx.kv.create("device")
    kv.init(5, mx.nd.zeros((E,), mx.gpu(0)))
    g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(3)]
    vals = [mx.nd.array(g[0], mx.gpu(1)), mx.nd.array(g[1], mx.cpu()), mx.nd.array(g[2], mx.gpu(0))]
    out = mx.nd.empty((E,), mx.gpu(1))
    kv.pushpull(5, vals, out=out)
    assert _bits_equal(out.asnumpy(), O.sum_device(g))
    host_out = mx.nd.empty((E,), mx.cpu())
    kv.pull(5, out=host_out)
    assert _bits_equal(host_out.asnumpy(), O.sum_device(g))


def test_more_outputs_than_one_launch_addresses():
    """40 outputs of one key (kMaxOut is 24): the surplus is served by copies after the launch"""
    _need(2)
    E = 4099
    rng = _rng(3)
    devs = list(range(min(mx.num_gpus(), 4)))
    kv = mx.kv.create("device")
    kv.init(1, mx.nd.zeros((E,), mx.gpu(0)))
    g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
    outs = [mx.nd.empty((E,), mx.gpu(devs[i % len(devs)])) for i in range(40)]
    kv.pushpull(1, [mx.nd.array(x, mx.gpu(d)) for x, d in zip(g, devs)], out=outs)
    want = O.sum_device(g)
    for i, o in enumerate(outs):
        assert _bits_equal(o.asnumpy(), want), i


@pytest.mark.parametrize("E", [1000, 300007])
def test_in_place_forms(E):
    """pushpull(key, values) overwrites the values with the result: one-shot keys must not be written while
    a peer still reads them (served by a copy afterwards), two-shot keys write disjoint shards directly"""
    _need(2)
    rng = _rng(4 + E)
    devs = list(range(min(mx.num_gpus(), 4)))
    kv = mx.kv.create("device")
    kv.init(["a", "b"], [mx.nd.zeros((E,), mx.gpu(0))] * 2)
    ga = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
    gb = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
    va = [mx.nd.array(x, mx.gpu(d)) for x, d in zip(ga, devs)]
    vb = [mx.nd.array(x, mx.gpu(d)) for x, d in zip(gb, devs)]
    kv.pushpull(["a", "b"], [va, vb])
    for v in va:
        assert _bits_equal(v.asnumpy(), O.sum_device(ga))
    for v in vb:
        assert _bits_equal(v.asnumpy(), O.sum_device(gb))
    # the same arrays pushed for BOTH keys and pulled in place (the reference's tests do this): every reduce of
    # the call happens before any pull (kvstore_local.h:358-365)
    shared = [mx.nd.array(x, mx.gpu(d)) for x, d in zip(ga, devs)]
    kv.pushpull(["a", "b"], [shared, shared])
    for v in shared:
        assert _bits_equal(v.asnumpy(), O.sum_device(ga))


def test_many_keys_many_gpus_layerwise_and_plain_mixed_dtypes():
    """fp32 and bf16 keys in one call (two launch classes), 8 GPUs if there are that many, LAMB with the
    device-side overflow skip spanning both classes"""
    _need(2)
    devs = list(range(min(mx.num_gpus(), 8)))
    rng = _rng(5)
    shapes = [(64,), (513, 9), (1 << 18,), (300, 1000)]
    kw = dict(learning_rate=0.01, wd=0.01)
    kv = mx.kv.create("device")
    w32 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    wbf = [O.f32_to_bf16(rng.uniform(-1, 1, s).astype(np.float32)) for s in shapes]
    keys32 = ["f%d" % i for i in range(len(shapes))]
    keysbf = ["h%d" % i for i in range(len(shapes))]
    kv.init(keys32, [mx.nd.array(w, mx.gpu(0)) for w in w32])
    kv.init(keysbf, [mx.nd.array(w, mx.gpu(0), dtype="bfloat16") for w in wbf])
    kv.set_optimizer(mx.optimizer.LAMB(skip_nonfinite=True, multi_precision=True, **kw))
    o32 = [[mx.nd.empty(s, mx.gpu(d)) for d in devs] for s in shapes]
    obf = [[mx.nd.empty(s, mx.gpu(d), dtype="bfloat16") for d in devs] for s in shapes]
    oopt = O.OracleOptimizer("lamb", norm_mode="f64", **kw)
    ref32 = [w.copy() for w in w32]
    for step in range(3):
        g32 = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in devs] for s in shapes]
        gbf = [[O.f32_to_bf16(rng.uniform(-1, 1, s).astype(np.float32)) for _ in devs] for s in shapes]
        overflow = step == 1
        if overflow:
            gbf[0][-1][3] = O.f32_to_bf16(np.array([np.inf], np.float32))[0]     # a bf16 key poisons the push
        kv.pushpull(keys32 + keysbf,
                    [[mx.nd.array(g, mx.gpu(d)) for g, d in zip(gs, devs)] for gs in g32] +
                    [[mx.nd.array(g, mx.gpu(d), dtype="bfloat16") for g, d in zip(gs, devs)] for gs in gbf],
                    out=o32 + obf)
        assert kv.overflow() == overflow
        for k in range(len(shapes)):
            if not overflow:
                oopt.update(k, ref32[k], O.sum_device(g32[k]).reshape(shapes[k]))
            first = o32[k][0].asnumpy()
            np.testing.assert_allclose(first, ref32[k], rtol=2e-6, atol=2e-7, err_msg=str((step, k)))
            for o in o32[k][1:]:
                assert _bits_equal(o.asnumpy(), first)
            hb = obf[k][0].asnumpy(raw=True)
            for o in obf[k][1:]:
                assert np.array_equal(o.asnumpy(raw=True), hb)
