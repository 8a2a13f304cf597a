"""Threading contract of the C ABI (callable from any thread; results ordered per array): several
Python threads drive one shared store (different keys) and a store of their own at the same time -- ctypes
releases the GIL, so the calls really overlap -- and every thread checks its own results bit for bit.  (Sorted last
on purpose: it exercises the locks, not the arithmetic.)"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_calls_from_several_threads():
    E, nthr, ngpu = 5000, 4, max(1, min(mx.num_gpus(), 2))
    shared = mx.kv.create("device")
    shared.init(list(range(nthr)), [mx.nd.zeros((E,), mx.gpu(0)) for _ in range(nthr)])
    kw = dict(learning_rate=0.1, momentum=0.9)
    shared.set_optimizer(mx.optimizer.SGD(**kw))
    errs = []

    def work(t):
        try:
            rng = np.random.default_rng(t)
            oopt = O.OracleOptimizer("sgd", **kw)
            ow = np.zeros(E, np.float32)
            own = mx.kv.create("device")
            own.init("k", mx.nd.zeros((E,), mx.gpu(t % ngpu)))
            for it in range(40):
                gs = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(2)]
                out = mx.nd.empty((E,), mx.gpu(0))
                shared.pushpull(t, [mx.nd.array(g, mx.gpu(d % ngpu)) for d, g in enumerate(gs)], out=out)
                oopt.update(t, ow, O.sum_device(gs))
                assert _bits_equal(out.asnumpy(), ow), ("shared store", t, it)
                o2 = mx.nd.empty((E,), mx.gpu(0))
                own.pushpull("k", [mx.nd.array(g, mx.gpu(0)) for g in gs], out=o2)
                assert _bits_equal(o2.asnumpy(), O.sum_device(gs)), ("own store", t, it)
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            errs.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(nthr)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs[:2]
