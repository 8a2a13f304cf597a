"""Trainer (python/mxnet/gluon/trainer.py driver loop) over torch parameters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))


def _model():
    import torch
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 10)).cuda()


@pytest.mark.parametrize("batched", [True, False])
def test_trainer_step_matches_oracle_sgd(batched):
    """update_on_kvstore (default): step = rescale 1/batch, per-parameter pushpull, fused SGD-momentum."""
    import torch
    model = _model()
    params = list(model.parameters())
    kw = dict(learning_rate=0.1, momentum=0.9, wd=1e-4)
    trainer = mx.Trainer(params, "sgd", dict(kw), kvstore="device", batched=batched)
    ref_w = [p.detach().cpu().numpy().copy() for p in params]
    opt = O.OracleOptimizer("sgd", **kw)
    x = torch.randn(32, 64, device="cuda")
    for step in range(3):
        model.zero_grad()
        model(x).square().mean().backward()
        grads = [p.grad.detach().cpu().numpy().copy() for p in params]
        trainer.step(32)
        torch.cuda.synchronize()
        opt.rescale_grad = 1.0 / 32
        for i, (w, g) in enumerate(zip(ref_w, grads)):
            opt.update(i, w.reshape(-1), g.reshape(-1))
        for p, w in zip(params, ref_w):
            assert _bits_equal(p.detach().cpu().numpy(), w), step
    assert trainer._update_on_kvstore is True


def test_trainer_decision_table_and_local_update():
    """tests/python/unittest/test_gluon_trainer.py:247-281 (single-machine rows): kvstore=None or
    update_on_kvstore=False -> updates run per device outside the store."""
    import torch
    model = _model()
    params = list(model.parameters())
    t = mx.Trainer(params, "sgd", {"learning_rate": 0.1}, kvstore=None)
    x = torch.randn(8, 64, device="cuda")
    model(x).sum().backward()
    before = [p.detach().clone() for p in params]
    t.step(8)
    assert t._kvstore is None and t._update_on_kvstore is False
    for p, b in zip(params, before):
        assert torch.allclose(p, b - 0.1 * p.grad / 8, rtol=1e-5, atol=1e-6)
    t2 = mx.Trainer(list(_model().parameters()), "sgd", {"learning_rate": 0.1}, kvstore="device",
                    update_on_kvstore=False)
    m2 = t2._params
    model2 = _model()
    t2 = mx.Trainer(list(model2.parameters()), "sgd", {"learning_rate": 0.1}, kvstore="device",
                    update_on_kvstore=False)
    model2(x).sum().backward()
    t2.step(8)
    assert t2._update_on_kvstore is False and t2._kvstore is not None
    with pytest.raises(ValueError):
        class NoOpt(mx.kv.KVStore):
            @staticmethod
            def is_capable(c):
                return False
        mx.Trainer(list(_model().parameters()), "sgd", {}, kvstore=NoOpt("device"), update_on_kvstore=True)._init_kvstore()


def test_trainer_save_load_states(tmp_path):
    import torch
    x = torch.randn(16, 64, device="cuda")

    def run(trainer, model, n):
        for _ in range(n):
            model.zero_grad()
            model(x).square().mean().backward()
            trainer.step(16)
        torch.cuda.synchronize()
        return [p.detach().cpu().numpy().copy() for p in model.parameters()]

    m1 = _model()
    t1 = mx.Trainer(list(m1.parameters()), "adam", {"learning_rate": 0.01}, kvstore="device")
    run(t1, m1, 2)
    f = str(tmp_path / "trainer.states")
    t1.save_states(f)
    mid = [p.detach().clone() for p in m1.parameters()]
    want = run(t1, m1, 2)
    m2 = _model()
    with torch.no_grad():
        for p, v in zip(m2.parameters(), mid):
            p.copy_(v)
    t2 = mx.Trainer(list(m2.parameters()), "adam", {"learning_rate": 0.01}, kvstore="device")
    t2.load_states(f)
    got = run(t2, m2, 2)
    for a, b in zip(got, want):
        assert _bits_equal(a, b)


def test_amp_loss_scaler_skips_local_update_on_overflow():
    """amp.init_trainer + scale_loss (amp.py:290-298,374-400) with parameters updated outside the
    kvstore: an inf gradient skips the whole update and halves the scale; clean steps are unscaled by
    rescale_grad (gluon/trainer.py:445-448, loss_scaler.py:44-79)."""
    import torch
    model = _model()
    params = list(model.parameters())
    t = mx.Trainer(params, "sgd", {"learning_rate": 0.1}, kvstore=None)
    mx.amp.init_trainer(t)
    assert t._amp_loss_scaler.loss_scale == 2. ** 16
    x = torch.randn(8, 64, device="cuda")
    # clean step
    model.zero_grad()
    with mx.amp.scale_loss(model(x).sum(), t) as scaled:
        scaled.backward()
    before = [p.detach().clone() for p in params]
    grads = [p.grad.detach().clone() / 2. ** 16 for p in params]
    t.step(8)
    for p, b, g in zip(params, before, grads):
        assert torch.allclose(p, b - 0.1 * g / 8, rtol=1e-5, atol=1e-6)
    # overflow
    model.zero_grad()
    with mx.amp.scale_loss(model(x).sum(), t) as scaled:
        scaled.backward()
    params[1].grad[0] = float("inf")
    before = [p.detach().clone() for p in params]
    t.step(8)
    for p, b in zip(params, before):
        assert torch.equal(p, b)
    assert t._amp_loss_scaler._next_loss_scale == 2. ** 15
    # the next step runs with the halved scale
    model.zero_grad()
    with mx.amp.scale_loss(model(x).sum(), t) as scaled:
        scaled.backward()
    assert t._amp_loss_scaler.loss_scale == 2. ** 16      # switched at the next has_overflow (loss_scaler.py:67)
    t.step(8)
    assert t._amp_loss_scaler.loss_scale == 2. ** 15


def test_amp_with_update_on_kvstore_and_device_side_skip():
    """LAMB created with skip_nonfinite=True, parameters updated on the kvstore: the push itself skips
    on overflow and the Trainer feeds KVStore.overflow() to the loss scaler."""
    import torch
    model = _model()
    params = list(model.parameters())
    t = mx.Trainer(params, mx.optimizer.LAMB(learning_rate=0.01, skip_nonfinite=True), kvstore="device")
    mx.amp.init_trainer(t)
    x = torch.randn(8, 64, device="cuda")
    model.zero_grad()
    with mx.amp.scale_loss(model(x).square().mean(), t) as scaled:
        scaled.backward()
    t.step(1)
    torch.cuda.synchronize()
    after_clean = [p.detach().clone() for p in params]
    model.zero_grad()
    with mx.amp.scale_loss(model(x).square().mean(), t) as scaled:
        scaled.backward()
    params[0].grad.view(-1)[5] = float("nan")
    t.step(1)
    torch.cuda.synchronize()
    for p, b in zip(params, after_clean):
        assert torch.equal(p, b)
    assert t._amp_loss_scaler._next_loss_scale == 2. ** 15


@pytest.mark.parametrize("optname,kw", [("sgd", dict(learning_rate=0.1, momentum=0.9, wd=1e-4)),
                                        ("adam", dict(learning_rate=0.01, wd=1e-3))])
def test_overlap_buckets_match_the_plain_step(optname, kw):
    """overlap=True: buckets of parameters are exchanged + updated from grad-ready hooks while backward is still
    running (priority = -bucket), `step` only fences.  Same bits as the plain step, every step, including a
    change of batch size announced with arm()."""
    import torch

    def net():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),
                                   torch.nn.ReLU(), torch.nn.Linear(256, 10)).cuda()
    ma, mb = net(), net()
    ta = mx.Trainer(list(ma.parameters()), optname, dict(kw), kvstore="device")
    tb = mx.Trainer(list(mb.parameters()), optname, dict(kw), kvstore="device", overlap=True, bucket_bytes=64 << 10)
    tb.trace = []
    torch.manual_seed(9)
    for step, bs in enumerate([32, 32, 32, 16, 16]):
        x = torch.randn(bs, 64, device="cuda")
        for m, t in ((ma, ta), (mb, tb)):
            m.zero_grad(set_to_none=False)
            if t is tb and step > 0:
                t.arm(bs)
            m(x).square().mean().backward()
            t.step(bs)
        torch.cuda.synchronize()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert _bits_equal(pa.detach().cpu().numpy(), pb.detach().cpu().numpy()), (optname, step)
    assert tb._buckets is not None and len(tb._buckets) >= 2
    # every overlapped step fired every bucket from a hook (4 overlapped steps)
    assert len(tb.trace) == 4 * len(tb._buckets)
    # a step() with a batch size the backward was not armed for is refused
    mb.zero_grad(set_to_none=False)
    mb(torch.randn(8, 64, device="cuda")).square().mean().backward()
    with pytest.raises(AssertionError):
        tb.step(8)
