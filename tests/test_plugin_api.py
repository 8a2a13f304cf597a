"""The plug-in side of the boundary: `KVStoreBase.register` + `create(name)` (python/mxnet/kvstore/base.py:74-243,
406-461), driven the way tests/python/unittest/test_kvstore_custom.py:148-184 drives the reference's own sample
plug-in.  The sample store here is written by the *test*, as a third party would write one, against this package's
base class; host arrays only, so no GPU is needed."""
import numpy as np
import pytest

import mxnet_b200 as mx
from mxnet_b200.kvstore import KVStoreBase


@KVStoreBase.register
class SampleStore(KVStoreBase):
    """broadcast copies, pushpull sums on the first value's context -- what the reference's sample plug-in does"""

    def broadcast(self, key, value, out, priority=0):
        for o in (out if isinstance(out, list) else [out]):
            o[:] = value

    def pushpull(self, key, value, out=None, priority=0):
        if isinstance(value, mx.nd.NDArray):
            if out is not None:
                for o in (out if isinstance(out, list) else [out]):
                    o[:] = value
            return
        ctx = value[0].context
        reduced = sum([v.as_in_context(ctx) for v in value])
        for o in (value if out is None else (out if isinstance(out, list) else [out])):
            o[:] = reduced

    @staticmethod
    def is_capable(capability):
        if capability.lower() == KVStoreBase.OPTIMIZER:
            return False
        raise ValueError("Unknown capability: {}".format(capability))

    @property
    def type(self):
        return "samplestore"

    @property
    def rank(self):
        return 0

    @property
    def num_workers(self):
        return 1


def check_diff_to_scalar(A, x):
    assert np.sum(np.abs(A.asnumpy() - x)) == 0, (A.asnumpy(), x)


def test_custom_store():
    # test_kvstore_custom.py:148-163
    kv = mx.kv.create("samplestore")
    assert isinstance(kv, SampleStore)
    out = mx.nd.empty((1,))
    kv.broadcast(1, mx.nd.ones((1,)), out=out)
    check_diff_to_scalar(out, 1)
    assert type(kv).is_capable("optimizer") is False
    kv.broadcast(1, mx.nd.ones((1,)), out=out)
    check_diff_to_scalar(out, 1)
    arr_list = [mx.nd.empty((1,))] * 2
    kv.pushpull(1, [mx.nd.ones((1,))] * 2, out=arr_list)
    for arr in arr_list:
        check_diff_to_scalar(arr, 2)
    kv.pushpull(1, arr_list)
    for arr in arr_list:
        check_diff_to_scalar(arr, 4)


def test_get_type_of_a_registered_store():
    # test_kvstore_custom.py:165-168; the name is the lower-cased class name, matched case-insensitively
    assert mx.kv.create("samplestore").type == "samplestore"
    assert isinstance(mx.kv.create("SampleStore"), SampleStore)


def test_set_optimizer_unsupported_by_a_plug_in():
    # test_kvstore_custom.py:170-180: what a plug-in does not override raises NotImplementedError
    kv = mx.kv.create("samplestore")
    assert not kv.is_capable("optimizer")
    with pytest.raises(NotImplementedError):
        kv.set_optimizer(mx.optimizer.create("sgd"))
    with pytest.raises(NotImplementedError):
        kv.save_optimizer_states("test")
    with pytest.raises(NotImplementedError):
        kv.load_optimizer_states("test")


def test_registry_rules():
    # base.py:225-243 (register takes classes), :432-435 (create takes a string); names that are not registered
    # fall through to the native factory, which knows the reference's built-in types and rejects the rest
    with pytest.raises(AssertionError):
        KVStoreBase.register(SampleStore())
    with pytest.raises(TypeError):
        mx.kv.create(3)
    # src/kvstore/kvstore.cc:42-84: any other name is a local store that reports the name it was given
    assert mx.kv.create("no_such_store").type == "no_such_store"
    with pytest.raises(mx.MXNetError):
        mx.kv.create("dist_sync")                            # out of scope of this library, said loudly
    assert "b200device" in KVStoreBase.kv_registry          # the engine's own entry
    assert mx.kv.KVStore.is_capable("optimizer") is True
    with pytest.raises(mx.MXNetError):
        mx.kv.KVStore.is_capable("no_such_capability")


def test_trainer_refuses_update_on_kvstore_for_a_plug_in_without_optimizer_support():
    # gluon/trainer.py:225-229: a store that is not capable of 'optimizer' cannot own the update;
    # with update_on_kvstore left to the trainer it falls back to local updates
    # (tests/nightly/dist_device_sync_kvstore.py:107-125 checks the same table for its store)
    class Param(object):
        def __init__(self, ctx):
            self.data = mx.nd.zeros((10, 1), ctx)
            self.grad = mx.nd.ones((10, 1), ctx)

    params = [[Param(mx.Context("cpu", 0)), Param(mx.Context("cpu", 1))]]
    kv = mx.kv.create("samplestore")
    tr = mx.Trainer(params, "sgd", {"learning_rate": 0.1}, kvstore=kv, update_on_kvstore=True)
    with pytest.raises(ValueError):
        tr._init_kvstore()
    for uok in (False, None):
        tr = mx.Trainer(params, "sgd", {"learning_rate": 0.1}, kvstore=kv, update_on_kvstore=uok)
        tr._init_kvstore()
        assert tr._kv_initialized and tr._update_on_kvstore is False
        # the plug-in is driven one key per call, as the reference's Trainer drives it
        tr.allreduce_grads()
        for p in params[0]:
            check_diff_to_scalar(p.grad, 2)
        for p in params[0]:
            p.grad[:] = 1
