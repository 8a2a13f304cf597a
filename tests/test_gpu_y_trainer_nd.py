"""Trainer (python/mxnet/gluon/trainer.py driver loop) over parameters held in engine NDArrays -- the
reference's own parameter type -- so that the loop can run without torch (e.g. on the simulator): the
decision table, update on the kvstore vs. per-device updaters, several replicas per parameter, AMP."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O


class Param(object):
    """What Trainer needs of a parameter replica: .data and .grad"""

    def __init__(self, w, ctx):
        self.data = mx.nd.array(w, ctx)
        self.grad = mx.nd.zeros(w.shape, ctx)


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


SHAPES = [(64, 33), (129,), (1 << 17,), (7,)]


@pytest.mark.parametrize("update_on_kvstore", [True, False])
@pytest.mark.parametrize("optname,kw", [
    ("sgd", dict(learning_rate=0.1, momentum=0.9, wd=1e-4)),
    ("adam", dict(learning_rate=0.01, wd=1e-3)),
])
def test_step_matches_oracle_over_replicas(update_on_kvstore, optname, kw):
    """trainer.py:334-480 with one replica of every parameter per GPU: rescale 1/batch, allreduce (+ update on
    the store, or per-device native updaters), every replica ends up bit-identical to the oracle"""
    ndev = min(mx.num_gpus(), 4)
    rng = np.random.default_rng(5)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES]
    params = [[Param(w, mx.gpu(d)) for d in range(ndev)] for w in w0]
    tr = mx.Trainer(params, optname, dict(kw), kvstore="device", update_on_kvstore=update_on_kvstore)
    oopt = O.OracleOptimizer(optname, **kw)
    ow = [w.copy() for w in w0]
    batch = 16
    for step in range(3):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(ndev)] for s in SHAPES]
        for reps, gs in zip(params, grads):
            for p, g in zip(reps, gs):
                p.grad[:] = g
        tr.step(batch)
        oopt.rescale_grad = 1.0 / batch
        for i in range(len(SHAPES)):
            oopt.update(i, ow[i], O.sum_device(grads[i]).reshape(SHAPES[i]) if ndev > 1 else grads[i][0])
            for p in params[i]:
                assert _bits_equal(p.data.asnumpy(), ow[i]), (optname, update_on_kvstore, step, i)
    assert tr._update_on_kvstore is update_on_kvstore


def test_kvstore_none_uses_local_updaters():
    rng = np.random.default_rng(6)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES]
    params = [Param(w, mx.gpu(0)) for w in w0]
    kw = dict(learning_rate=0.1, momentum=0.9)
    tr = mx.Trainer(params, "sgd", dict(kw), kvstore=None)
    oopt = O.OracleOptimizer("sgd", **kw)
    ow = [w.copy() for w in w0]
    for step in range(2):
        for p, s in zip(params, SHAPES):
            p.grad[:] = rng.uniform(-1, 1, s).astype(np.float32)
        gs = [p.grad.asnumpy().copy() for p in params]
        tr.step(4)
        oopt.rescale_grad = 0.25
        for i, p in enumerate(params):
            oopt.update(i, ow[i], gs[i])
            assert _bits_equal(p.data.asnumpy(), ow[i])
    assert tr._kvstore is None and isinstance(tr._updaters[0], mx.optimizer.NativeUpdater)


def test_amp_loss_scaling_both_ways():
    """dynamic loss scaling: skip in the Trainer (local updaters) and skip on the device (LAMB, update on kvstore)"""
    rng = np.random.default_rng(7)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES[:2]]
    for on_kv in (False, True):
        params = [Param(w, mx.gpu(0)) for w in w0]
        if on_kv:
            tr = mx.Trainer(params, mx.optimizer.LAMB(learning_rate=0.01, skip_nonfinite=True), kvstore="device")
        else:
            tr = mx.Trainer(params, "sgd", {"learning_rate": 0.1}, kvstore=None)
        mx.amp.init_trainer(tr)
        scale = tr._amp_loss_scaler.loss_scale

        def backward(poison=False):
            tr._scale = tr._amp_original_scale / tr._amp_loss_scaler.loss_scale       # what amp.scale_loss does
            for p, s in zip(params, SHAPES):
                g = rng.uniform(-1, 1, s).astype(np.float32) * np.float32(tr._amp_loss_scaler.loss_scale)
                if poison and s == SHAPES[1]:
                    g[3] = np.inf
                p.grad[:] = g
        backward()
        tr.step(1)
        after = [p.data.asnumpy().copy() for p in params]
        assert not any(_bits_equal(a, w) for a, w in zip(after, w0))
        backward(poison=True)
        tr.step(1)
        for p, a in zip(params, after):
            assert _bits_equal(p.data.asnumpy(), a), on_kv
        assert tr._amp_loss_scaler._next_loss_scale == scale / 2
        backward()
        tr.step(1)
        assert tr._amp_loss_scaler.loss_scale == scale / 2
        assert not any(_bits_equal(p.data.asnumpy(), a) for p, a in zip(params, after))


def test_save_load_states_round_trip(tmp_path):
    rng = np.random.default_rng(8)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES]
    grads = [[rng.uniform(-1, 1, s).astype(np.float32) for s in SHAPES] for _ in range(4)]

    def run(tr, params, gs_list):
        for gs in gs_list:
            for p, g in zip(params, gs):
                p.grad[:] = g
            tr.step(2)

    for on_kv in (True, False):
        pa = [Param(w, mx.gpu(0)) for w in w0]
        ta = mx.Trainer(pa, "adam", {"learning_rate": 0.01}, kvstore="device", update_on_kvstore=on_kv)
        run(ta, pa, grads[:2])
        f = str(tmp_path / ("t_%d.states" % on_kv))
        ta.save_states(f)
        mid = [p.data.asnumpy().copy() for p in pa]
        run(ta, pa, grads[2:])
        pb = [Param(w, mx.gpu(0)) for w in mid]
        tb = mx.Trainer(pb, "adam", {"learning_rate": 0.01}, kvstore="device", update_on_kvstore=on_kv)
        tb.load_states(f)
        run(tb, pb, grads[2:])
        for a, b in zip(pa, pb):
            assert _bits_equal(a.data.asnumpy(), b.data.asnumpy()), on_kv


class RefParam(Param):
    """a gluon.Parameter as far as the optimizer is concerned: replicas + lr_mult / wd_mult read at every update"""

    def __init__(self, w, ctx):
        super(RefParam, self).__init__(w, ctx)
        self.lr_mult = 1.0
        self.wd_mult = 1.0


def _two_replicas(shape=(10,)):
    ndev = max(1, min(mx.num_gpus(), 2))
    reps = [RefParam(np.zeros(shape, np.float32), mx.gpu(d % ndev)) for d in range(2)]
    reps[1].__dict__.pop("lr_mult"), reps[1].__dict__.pop("wd_mult")        # the Parameter is replica 0's owner
    return reps


def _backward_of_w_plus_1(reps):
    for p in reps:
        p.grad[:] = 1           # d(w + 1)/dw


@pytest.mark.parametrize("update_on_kvstore", [None, False])
def test_reference_trainer_known_answers(update_on_kvstore, tmp_path):
    # tests/python/unittest/test_gluon_trainer.py:80-128: two replicas, sgd lr 1 momentum 0.5; every replica
    # contributes a gradient of ones -> -2 after one step; then lr_mult = 0.5 ON THE PARAMETER -> -4
    x = _two_replicas()
    trainer = mx.Trainer([x], "sgd", {"learning_rate": 1.0, "momentum": 0.5}, update_on_kvstore=update_on_kvstore)
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert (x[1].data.asnumpy() == -2).all()
    x[0].lr_mult = 0.5
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert (x[1].data.asnumpy() == -4).all(), x[1].data.asnumpy()
    assert (x[0].data.asnumpy() == -4).all()

    f = str(tmp_path / "test_trainer.states")
    trainer.save_states(f)
    trainer.load_states(f)
    if trainer._update_on_kvstore:
        assert trainer._optimizer is trainer._kvstore._optimizer
        # invalid usage of update and allreduce_grads if update_on_kvstore (:109-111)
        with pytest.raises(AssertionError):
            trainer.update(1)
        with pytest.raises(AssertionError):
            trainer.allreduce_grads()
    else:
        assert all(u.optimizer is trainer._optimizer for u in trainer._updaters)
    # the loaded optimizer carries on: momentum -2 -> 0.5 * -2 - 0.5 * 2 = -2 -> w = -6
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert (x[0].data.asnumpy() == -6).all(), x[0].data.asnumpy()


def test_reference_trainer_allreduce_then_update():
    # test_gluon_trainer.py:117-129: gradients that differ per replica (i * w -> i), allreduce_grads makes them
    # equal, update(1) applies them: 0 - 1 * (0 + 1) = -1
    x = _two_replicas()
    trainer2 = mx.Trainer([x], "sgd", {"learning_rate": 1.0, "momentum": 0.5}, update_on_kvstore=False)
    for i, p in enumerate(x):
        p.grad[:] = float(i)
    assert (x[0].grad.asnumpy() != x[1].grad.asnumpy()).all()
    trainer2.allreduce_grads()
    assert (x[0].grad.asnumpy() == x[1].grad.asnumpy()).all()
    trainer2.update(1)
    assert (x[1].data.asnumpy() == -1).all(), x[1].data.asnumpy()


def test_reference_trainer_save_load_reattaches_parameters(tmp_path):
    # test_gluon_trainer.py:131-149: after load_states the optimizer reads the CURRENT parameters' multipliers
    x = _two_replicas()
    trainer = mx.Trainer([x], "sgd", {"learning_rate": 0.1})
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert trainer._kvstore._optimizer._get_lr(0) == 0.1
    f = str(tmp_path / "test_trainer_save_load.states")
    trainer.save_states(f)
    trainer.load_states(f)
    x[0].lr_mult = 2.0
    assert trainer._kvstore._optimizer._get_lr(0) == 0.2
    # ... and so does the engine: -0.2 - 0.2 * 2
    _backward_of_w_plus_1(x)
    trainer.step(1)
    np.testing.assert_array_equal(x[1].data.asnumpy(), np.float32(np.float32(-0.2) - np.float32(0.2) * np.float32(2)))


@pytest.mark.parametrize("update_on_kvstore", [None, False])
def test_reference_trainer_lr_sched(update_on_kvstore):
    # test_gluon_trainer.py:283-320: FactorScheduler(2, 0.1); with two replicas every step counts ONCE
    # (per-device updaters count per device id, optimizer.py:433-443: the replicas must sit on two devices)
    if update_on_kvstore is False and mx.num_gpus() < 2:
        pytest.skip("needs two GPUs")
    x = _two_replicas()
    freq, factor, lr = 2, 0.1, 1
    sched = mx.lr_scheduler.FactorScheduler(freq, factor=factor, base_lr=lr)
    trainer = mx.Trainer([x], "sgd", {"learning_rate": lr, "lr_scheduler": sched}, update_on_kvstore=update_on_kvstore)
    for i in range(10):
        _backward_of_w_plus_1(x)
        trainer.step(1)
        if i % freq == 0:
            assert trainer.learning_rate == lr, (lr, trainer.learning_rate, i)
            lr *= factor


class SparseParam(object):
    """Parameter(stype=..., grad_stype=...) replica: the arrays carry the storage types"""

    def __init__(self, shape, ctx, stype, grad_stype):
        self.data = mx.nd.zeros(shape, ctx, stype=stype)
        self.grad = mx.nd.zeros(shape, ctx, stype=grad_stype)
        self.lr_mult = 1.0
        self.wd_mult = 1.0

    def set_grad_ones(self):
        ones = mx.nd.ones(self.data.shape, self.grad.context)
        (ones.tostype("row_sparse") if self.grad.stype == "row_sparse" else ones).copyto(self.grad)


@pytest.mark.parametrize("kv", ["local", "device"])
@pytest.mark.parametrize("stype,grad_stype,update_on_kv,expected", [
    ("default", "default", True, True),
    ("default", "default", False, False),
    ("default", "default", None, True),
    ("default", "row_sparse", None, False),
    ("default", "row_sparse", True, True),
    ("default", "row_sparse", False, False),
    ("row_sparse", "row_sparse", None, True),
    ("row_sparse", "row_sparse", False, ValueError),
])
def test_reference_trainer_sparse_kv(kv, stype, grad_stype, update_on_kv, expected):
    # tests/python/unittest/test_gluon_trainer.py:247-281: the storage-type rows of the decision table; every
    # variant ends at w = 0 - 0.1 * (1 + 1) on all ten rows
    shape = (10, 1)
    ndev = max(1, min(mx.num_gpus(), 2))
    x = [SparseParam(shape, mx.gpu(d % ndev), stype, grad_stype) for d in range(2)]
    trainer = mx.Trainer([x], "sgd", {"learning_rate": 0.1}, kvstore=kv, update_on_kvstore=update_on_kv)
    all_rows = mx.nd.array(np.arange(10), dtype=np.int64)
    if expected is ValueError:
        with pytest.raises(ValueError):
            trainer.step(1)
        return
    if stype == "row_sparse":                                   # list_row_sparse_data(all_rows) before forward
        trainer._row_sparse_pull(0, [p.data for p in x], all_rows)
        for p in x:
            assert np.all(p.data.todense_numpy() == 0)
    for p in x:
        p.set_grad_ones()
    trainer.step(1)
    assert trainer._kvstore.type == kv
    assert trainer._kv_initialized
    assert trainer._update_on_kvstore is expected
    mx.nd.waitall()
    if stype == "default":
        for p in x:
            assert (p.data.asnumpy() == np.float32(-0.2)).all(), p.data.asnumpy()
    else:
        out = mx.nd.zeros(shape, mx.gpu(0), stype="row_sparse")
        trainer._row_sparse_pull(0, out, all_rows)
        assert (out.todense_numpy() == np.float32(-0.2)).all(), out.todense_numpy()
        out2 = mx.nd.zeros(shape, mx.gpu(0), stype="row_sparse")
        trainer._row_sparse_pull(x[0], out2, all_rows, full_idx=True)
        assert (out2.todense_numpy() == np.float32(-0.2)).all()


def test_reference_trainer_sparse_save_load(tmp_path):
    # test_gluon_trainer.py:151-167: row_sparse weight and gradient on one context, update on the store,
    # save / load, then the parameter's lr_mult is still consulted
    x = [SparseParam((10, 1), mx.gpu(0), "row_sparse", "row_sparse")]
    trainer = mx.Trainer([x], "sgd", {"learning_rate": 0.1})
    all_rows = mx.nd.array(np.arange(10), dtype=np.int64)
    trainer._row_sparse_pull(0, x[0].data, all_rows)
    x[0].set_grad_ones()
    trainer.step(1)
    assert trainer._kvstore._optimizer._get_lr(0) == 0.1
    f = str(tmp_path / "test_trainer_sparse_save_load.states")
    trainer.save_states(f)
    trainer.load_states(f)
    x[0].lr_mult = 2.0
    assert trainer._kvstore._optimizer._get_lr(0) == 0.2
    x[0].set_grad_ones()
    trainer.step(1)
    trainer._row_sparse_pull(0, x[0].data, all_rows)
    want = np.float32(np.float32(-0.1) - np.float32(0.2) * np.float32(1))
    assert (x[0].data.todense_numpy() == want).all(), x[0].data.todense_numpy()


def test_reference_trainer_sparse_grad_single_context():
    # test_gluon_trainer.py:48-60: 1-D parameter, row_sparse gradient, sgd momentum 0.5 lr 1 -> -1.  (The
    # reference creates no store for a single context; this engine keeps one -- see Trainer._init_kvstore --
    # and the result is the same.)
    x = [SparseParam((10,), mx.gpu(0), "default", "row_sparse")]
    trainer = mx.Trainer([x], "sgd", {"learning_rate": 1.0, "momentum": 0.5})
    x[0].set_grad_ones()
    trainer.step(1)
    assert trainer._update_on_kvstore is False
    assert (x[0].data.asnumpy() == -1).all(), x[0].data.asnumpy()


@pytest.mark.parametrize("kv", ["local", "device"])
def test_reference_trainer_reset_kv(kv):
    # test_gluon_trainer.py:213-245: parameters loaded from a checkpoint after the first step drop the store;
    # the next step creates it again from the loaded values: 0 - 0.1 * 2 = -0.2 both times
    x = _two_replicas((10, 1))
    trainer = mx.Trainer({"x": x}, "sgd", {"learning_rate": 0.1}, kvstore=kv)
    saved = x[0].data.asnumpy().copy()
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert trainer._kvstore.type == kv
    mx.nd.waitall()
    for p in x:                              # x._load_init(...) -> trainer._reset_kvstore()
        p.data[:] = saved
    trainer._reset_kvstore()
    assert trainer._kvstore is None and trainer._kv_initialized is False
    _backward_of_w_plus_1(x)
    trainer.step(1)
    assert (x[0].data.asnumpy() == np.float32(-0.2)).all(), x[0].data.asnumpy()
    assert (x[1].data.asnumpy() == np.float32(-0.2)).all()


def test_rescale_change_keeps_python_updater_state():
    # a smaller last batch changes rescale_grad; a Python (non-fused) optimizer on the store reads
    # optimizer.rescale_grad live through its Updater (updater.py:39-93) and must keep its states
    @mx.optimizer.register
    class CountingStep(mx.optimizer.Optimizer):
        def create_state(self, index, weight):
            return mx.nd.zeros(weight.shape, weight.context)

        def step(self, indices, weights, grads, states):
            self._update_count(indices)
            for i, w, g, s, lr in zip(indices, weights, grads, states, self._get_lrs(indices)):
                s[:] = s.asnumpy() + 1
                w[:] = w.asnumpy() - lr * self.rescale_grad * g.asnumpy() * s.asnumpy()

    shape = (4, 4)
    x = [Param(np.ones(shape, np.float32), mx.gpu(0))]
    tr = mx.Trainer([x], CountingStep(learning_rate=1.0), kvstore="device", update_on_kvstore=True)
    w = np.ones(shape, np.float32)
    for k, bs in enumerate((1, 1, 2), start=1):
        x[0].grad[:] = np.ones(shape, np.float32)
        tr.step(bs)
        w = w - np.float32(1.0 / bs) * np.float32(k)          # state = k: it survives the rescale change
        np.testing.assert_allclose(x[0].data.asnumpy(), w, rtol=1e-6)


def test_save_states_before_first_step(tmp_path):
    # with update_on_kvstore=False the per-device updaters exist from initialisation on
    x = [Param(np.ones((4, 4), np.float32), mx.gpu(0))]
    tr = mx.Trainer([x], "sgd", {"learning_rate": 0.1, "momentum": 0.9}, kvstore="device", update_on_kvstore=False)
    tr._init_kvstore()
    f = str(tmp_path / "t.states")
    tr.save_states(f)
    tr.load_states(f)
