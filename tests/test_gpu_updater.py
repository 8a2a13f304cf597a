"""The native per-device updater (MXKVStoreCreate("updater") + MXKVB200UpdaterStep): in-place fused
multi-tensor updates of caller-owned weights -- the reference's Updater -> multi_sgd_mom_update /
multi_mp_sgd_* / multi_adamw / multi_lamb path for parameters that are NOT updated on the kvstore
(python/mxnet/optimizer/updater.py:39-93, sgd.py:170-213) -- against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O

SHAPES = [(64,), (3, 5), (1000,), (257, 33), (2048, 160), (7,)]


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("name,kw,exact", [
    ("sgd", dict(learning_rate=0.1, wd=1e-4, rescale_grad=0.5), True),
    ("sgd", dict(learning_rate=0.1, momentum=0.9, wd=1e-4, clip_gradient=0.3), True),
    ("adam", dict(learning_rate=0.01, wd=1e-3), True),
    ("lamb", dict(learning_rate=0.01, wd=0.01), False),
    ("lars", dict(learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01), False),
])
def test_in_place_multi_tensor_update(name, kw, exact):
    rng = np.random.default_rng(5)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES]
    weights = [mx.nd.array(w, mx.gpu(0)) for w in w0]
    ptrs = [w.data_ptr for w in weights]
    upd = mx.optimizer.get_updater(mx.optimizer.create(name, **kw))
    assert isinstance(upd, mx.optimizer.NativeUpdater)
    okw = dict(kw)
    oopt = O.OracleOptimizer(name, norm_mode="f64", **okw) if not exact else O.OracleOptimizer(name, **okw)
    ow = [w.copy() for w in w0]
    idx = list(range(len(SHAPES)))
    for step in range(3):
        g = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES]
        before = mx.kv.launch_count()
        upd(idx, [mx.nd.array(x, mx.gpu(0)) for x in g], weights)
        launches = mx.kv.launch_count() - before
        assert launches == (1 if exact else 3), "one launch (sequence) for the whole parameter list"
        for k in idx:
            oopt.update(k, ow[k], g[k].copy())
            got = weights[k].asnumpy()
            if exact:
                assert _bits_equal(got, ow[k]), (name, step, k)
            else:
                np.testing.assert_allclose(got, ow[k], rtol=2e-6, atol=2e-7, err_msg=str((name, step, k)))
    assert [w.data_ptr for w in weights] == ptrs, "updated in place"
    assert upd.optimizer._index_update_count[0] == 3


def test_single_index_call_and_str_keys():
    rng = np.random.default_rng(6)
    w0 = rng.uniform(-1, 1, 5003).astype(np.float32)
    kw = dict(learning_rate=0.1, momentum=0.9)
    for key in (3, "fc1_weight"):
        w = mx.nd.array(w0, mx.gpu(0))
        upd = mx.optimizer.get_updater(mx.optimizer.SGD(**kw))
        oopt = O.OracleOptimizer("sgd", **kw)
        ow = w0.copy()
        for _ in range(2):
            g = rng.uniform(-1, 1, 5003).astype(np.float32)
            upd(key, mx.nd.array(g, mx.gpu(0)), w)               # updater.py:39: scalar index form
            oopt.update(0, ow, g)
            assert _bits_equal(w.asnumpy(), ow)


@pytest.mark.parametrize("lp", ["bfloat16", np.float16])
def test_multi_precision_in_place(lp):
    """multi_mp_sgd_mom_update: 16-bit weight updated in place, fp32 master and momentum in the updater."""
    kind = 2 if lp == "bfloat16" else 1
    to_lp = (lambda x: O.f32_to_bf16(x)) if kind == 2 else (lambda x: x.astype(np.float16))
    to_f32 = (lambda x: O.bf16_to_f32(x)) if kind == 2 else (lambda x: x.astype(np.float32))
    rng = np.random.default_rng(7)
    E = 30011
    w_lp = to_lp(rng.uniform(-1, 1, E).astype(np.float32))
    w32, mom = to_f32(w_lp), np.zeros(E, np.float32)
    w = mx.nd.array(w_lp, mx.gpu(0), dtype=lp)
    upd = mx.optimizer.get_updater(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=1e-4, multi_precision=True))
    want = np.zeros(E, np.uint16)
    for _ in range(3):
        g_lp = to_lp(rng.uniform(-1, 1, E).astype(np.float32))
        upd(0, mx.nd.array(g_lp, mx.gpu(0), dtype=lp), w)
        O.mp_sgd_mom_update(want, kind, w32, mom, to_f32(g_lp), 0.1, 1e-4, 0.9)
        got = w.asnumpy(raw=True) if kind == 2 else w.asnumpy().view(np.uint16)
        assert np.array_equal(got, want)


def test_states_round_trip():
    """Updater.get_states / set_states (updater.py:108-127): states loaded BEFORE the first call."""
    rng = np.random.default_rng(8)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES[:4]]
    grads = [[rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES[:4]] for _ in range(4)]
    idx = list(range(4))
    kw = dict(learning_rate=0.01, wd=1e-3)

    def run(upd, ws, gs_list):
        for gs in gs_list:
            upd(idx, [mx.nd.array(g, mx.gpu(0)) for g in gs], ws)

    wa = [mx.nd.array(w, mx.gpu(0)) for w in w0]
    ua = mx.optimizer.get_updater(mx.optimizer.Adam(**kw))
    run(ua, wa, grads[:2])
    blob = ua.get_states(dump_optimizer=True)         # the optimizer carries the update counts
    mid = [w.asnumpy().copy() for w in wa]
    run(ua, wa, grads[2:])
    wb = [mx.nd.array(w, mx.gpu(0)) for w in mid]
    ub = mx.optimizer.get_updater(mx.optimizer.Adam(**kw))
    ub.set_states(blob)
    run(ub, wb, grads[2:])
    for a, b in zip(wa, wb):
        assert _bits_equal(a.asnumpy(), b.asnumpy())


def test_torch_parameters_updated_in_place():
    """The arrays may be views of torch tensors (what Trainer binds): torch sees the new weights."""
    import torch
    t = torch.linspace(-1, 1, 4099, device="cuda")
    g = torch.full_like(t, 0.5)
    before = t.clone()
    upd = mx.optimizer.get_updater(mx.optimizer.SGD(learning_rate=0.1))
    upd(0, mx.nd.from_torch(g), mx.nd.from_torch(t))
    mx.nd.waitall()
    assert torch.allclose(t, before - 0.1 * g, rtol=0, atol=1e-7)


def test_rejects_misuse():
    kv = mx.kv.create("device")
    w = mx.nd.zeros((8,), mx.gpu(0))
    from mxnet_b200.base import _LIB, check_call, MXNetError
    import ctypes
    with pytest.raises(MXNetError):           # not an 'updater' store
        check_call(_LIB.MXKVB200UpdaterStep(kv.handle, 1, (ctypes.c_int * 1)(0), (ctypes.c_void_p * 1)(w.handle.value),
                                            (ctypes.c_void_p * 1)(w.handle.value)))
    upd = mx.optimizer.get_updater(mx.optimizer.SGD(learning_rate=0.1))
    with pytest.raises(MXNetError):           # gradient on the host
        upd(0, mx.nd.zeros((8,), mx.cpu()), w)
