"""GPU parity of the layer-wise adaptive optimizers (LAMB / LANS / LARS) fused with the gradient
exchange, and of the stand-alone multi_sum_sq / multi_all_finite reductions, against the CPU oracle
through the C ABI.

Tolerances.  The element arithmetic is the reference's, operation by operation; the only thing that
cannot be bit-identical is the rounding of the sums of squares (the reference's own CPU and GPU
operators already disagree there: sequential float sum vs block tree).  The oracle is therefore run
with the sums accumulated in double (``norm_mode='f64'``) and the comparison allows the few-ulp
effect of a ~1e-7 relative difference in a trust ratio: rtol 2e-6 on weights.  Whatever does not
depend on a norm (LAMB's mean / var) is compared bit for bit.  The reference's own tolerance for these
optimizers is rtol = atol = 1e-3 (tests/python/unittest/test_optimizer.py:251-312)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O

RTOL, ATOL = 2e-6, 2e-7


def _rng(seed):
    return np.random.default_rng(4321 + seed)


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _state(kv, key, which):
    h = kv._state_handle(key, which)
    return None if h is None else h.asnumpy()


def test_multi_sum_sq_matches_oracle():
    rng = _rng(1)
    sizes = [1, 7, 1000, 8192, 8193, 100003, (1 << 20) + 5]
    xs = [rng.uniform(-1, 1, n).astype(np.float32) for n in sizes]
    arrs = [mx.nd.array(x, mx.gpu(0)) for x in xs]
    for scale in (1.0, 0.25):
        got = mx.nd.multi_sum_sq(*arrs, scale=scale).asnumpy()
        want = np.array([O.sum_sq(x, scale, mode="f64") for x in xs], np.float32)
        np.testing.assert_allclose(got, want, rtol=1e-6)
    # run-to-run reproducible (fixed reduction shape)
    again = mx.nd.multi_sum_sq(*arrs).asnumpy()
    assert _bits_equal(again, mx.nd.multi_sum_sq(*arrs).asnumpy())
    # 16-bit inputs are squared in float32 (multi_sum_sq.cc:50: static_cast<float>)
    h = [x.astype(np.float16) for x in xs[:4]]
    got = mx.nd.multi_sum_sq(*[mx.nd.array(x, mx.gpu(0), dtype=np.float16) for x in h]).asnumpy()
    want = np.array([O.sum_sq(x.astype(np.float32), mode="f64") for x in h], np.float32)
    np.testing.assert_allclose(got, want, rtol=1e-6)
    b = [O.f32_to_bf16(x) for x in xs[:4]]
    got = mx.nd.multi_sum_sq(*[mx.nd.array(x, mx.gpu(0), dtype="bfloat16") for x in b]).asnumpy()
    want = np.array([O.sum_sq(O.bf16_to_f32(x), mode="f64") for x in b], np.float32)
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_multi_all_finite():
    rng = _rng(2)
    xs = [rng.uniform(-1, 1, n).astype(np.float32) for n in (5, 9000, 100003)]
    arrs = [mx.nd.array(x, mx.gpu(0)) for x in xs]
    assert mx.nd.multi_all_finite(*arrs).asnumpy()[0] == 1.0
    for bad in (np.inf, -np.inf, np.nan):
        y = xs[2].copy(); y[77777] = bad
        out = mx.nd.multi_all_finite(arrs[0], arrs[1], mx.nd.array(y, mx.gpu(0)))
        assert out.asnumpy()[0] == 0.0
        # init_output=False accumulates into an existing flag (all_finite.cu:52-53)
        mx.nd.multi_all_finite(*arrs, init_output=False, out=out)
        assert out.asnumpy()[0] == 0.0


def _run_case(name, kw, okw, n, E, steps, kvtype="device", exact_state=False):
    rng = _rng(E + 13 * n + steps)
    w0 = rng.uniform(-1, 1, E).astype(np.float32)
    kv = mx.kv.create(kvtype)
    kv.init(3, mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.create(name, **kw))
    okv = O.OracleKVStore(kvtype)
    okv.init(3, w0.copy())
    oopt = O.OracleOptimizer(name, norm_mode="f64", **okw)
    okv.set_optimizer(oopt)
    out = mx.nd.empty((E,), mx.gpu(0))
    oout = np.empty(E, np.float32)
    for s in range(steps):
        grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        if s % 2 == 0:
            kv.push(3, [mx.nd.array(g, mx.gpu(0)) for g in grads])
            kv.pull(3, out=out)
        else:
            kv.pushpull(3, [mx.nd.array(g, mx.gpu(0)) for g in grads], out=out)
        okv.push(3, [g.copy() for g in grads])
        okv.pull(3, oout)
        np.testing.assert_allclose(out.asnumpy(), oout, rtol=RTOL, atol=ATOL,
                                   err_msg="%s step %d n=%d E=%d" % (name, s, n, E))
        if exact_state:
            mean, var = oopt.states[3]
            assert _bits_equal(_state(kv, 3, 2), mean), "%s mean step %d" % (name, s)
            assert _bits_equal(_state(kv, 3, 3), var), "%s var step %d" % (name, s)
    return kv, oopt


@pytest.mark.parametrize("n", [1, 2, 5])
@pytest.mark.parametrize("bias_correction", [True, False])
def test_fused_lamb(n, bias_correction):
    kw = dict(learning_rate=0.01, wd=0.03, beta1=0.9, beta2=0.999, epsilon=1e-6, bias_correction=bias_correction)
    for E in (9, 4099, 300007):
        _run_case("lamb", kw, kw, n, E, 4, exact_state=True)


def test_fused_lamb_options():
    """clip, rescale, bounds, beta overrides (the option grid of test_optimizer.py:257-283)."""
    kw = dict(learning_rate=0.02, wd=0.03, beta1=0.5, beta2=0.8, epsilon=1e-6, clip_gradient=0.4, rescale_grad=0.14,
              lower_bound=1e-3, upper_bound=10.0)
    _run_case("lamb", kw, kw, 3, 50003, 3, exact_state=True)
    # upper bound that actually binds: ||w|| ~ sqrt(E/3) >> 1
    kw = dict(learning_rate=0.02, upper_bound=1.0)
    _run_case("lamb", kw, kw, 2, 20011, 2, exact_state=True)
    # CommCPU association order for kv.create('local')
    kw = dict(learning_rate=0.01, wd=0.01)
    _run_case("lamb", kw, kw, 6, 4099, 2, kvtype="local", exact_state=True)


@pytest.mark.parametrize("n", [1, 3])
def test_fused_lans(n):
    kw = dict(learning_rate=0.01, wd=0.03, beta1=0.9, beta2=0.999, epsilon=1e-6)
    for E in (9, 4099, 300007):
        _run_case("lans", kw, kw, n, E, 4)
    kw = dict(learning_rate=0.02, wd=0.03, beta1=0.5, beta2=0.8, epsilon=1e-6, clip_gradient=0.4, rescale_grad=0.14,
              lower_bound=1e-3, upper_bound=10.0)
    _run_case("lans", kw, kw, 2, 50003, 3)


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_fused_lars(momentum):
    kw = dict(learning_rate=0.1, momentum=momentum, wd=0.05, eta=0.01, epsilon=1e-8)
    for E in (9, 4099, 300007):
        _run_case("lars", kw, kw, 2, E, 4)
    kw = dict(learning_rate=0.1, momentum=momentum, wd=0.03, eta=0.002, clip_gradient=0.4, rescale_grad=0.14)
    _run_case("lars", dict(kw, epsilon=1e-8), kw, 4, 50003, 3)


def test_lars_skips_ratio_for_gamma_beta_bias():
    """lars.py:121-123: names ending in gamma / beta / bias keep the plain learning rate -- there the
    update is exactly sgd's and must be bit-identical."""
    E = 5003
    rng = _rng(7)
    w = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(2)]
    names = {0: "fc1_weight", 1: "bn1_gamma"}
    kv = mx.kv.create("device")
    kv.init([0, 1], [mx.nd.array(x, mx.gpu(0)) for x in w])
    kv.set_optimizer(mx.optimizer.LARS(learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01, param_idx2name=names))
    oopt = O.OracleOptimizer("lars", learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01, no_trust=[1], norm_mode="f64")
    outs = [mx.nd.empty((E,), mx.gpu(0)) for _ in range(2)]
    ow = [x.copy() for x in w]
    for s in range(3):
        g = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(2)]
        kv.pushpull([0, 1], [mx.nd.array(x, mx.gpu(0)) for x in g], out=outs)
        for k in range(2):
            oopt.update(k, ow[k], g[k].copy())
        np.testing.assert_allclose(outs[0].asnumpy(), ow[0], rtol=RTOL, atol=ATOL)
        assert _bits_equal(outs[1].asnumpy(), ow[1]), "plain sgd-momentum key step %d" % s


def test_multi_key_mixed_sizes_one_sequence():
    """a list of keys (small and large) goes through ONE first/finalize/apply sequence."""
    shapes = [(64,), (3, 5), (1000,), (257, 33), (2048, 160), (7,)]
    n = 3
    rng = _rng(9)
    keys = list(range(len(shapes)))
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    kw = dict(learning_rate=0.01, wd=0.01)
    kv.set_optimizer(mx.optimizer.LAMB(**kw))
    oopt = O.OracleOptimizer("lamb", norm_mode="f64", **kw)
    ow = [w.copy() for w in w0]
    outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
    for step in range(2):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(n)] for s in shapes]
        before = mx.kv.launch_count()
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(0)) for g in gs] for gs in grads], out=outs)
        assert mx.kv.launch_count() - before == 3, "first + finalize + apply"
        for k in keys:
            oopt.update(k, ow[k], O.sum_device(grads[k]).reshape(shapes[k]))
            np.testing.assert_allclose(outs[k].asnumpy(), ow[k], rtol=RTOL, atol=ATOL, err_msg="key %d" % k)


@pytest.mark.parametrize("lp", ["bfloat16", np.float16])
@pytest.mark.parametrize("name", ["lamb", "lans", "lars"])
def test_multi_precision(name, lp):
    """16-bit weights and gradients, fp32 master / mean / var (multi_mp_lamb_update etc.).  LAMB takes
    r1 from the master, LANS and LARS from the stored 16-bit weight (multi_lans-inl.h:296-300, lars.py:119)."""
    kind = 2 if lp == "bfloat16" else 1
    E, n = 50003, 2
    rng = _rng(17)
    to_lp = (lambda x: O.f32_to_bf16(x)) if kind == 2 else (lambda x: x.astype(np.float16))
    to_f32 = (lambda x: O.bf16_to_f32(x)) if kind == 2 else (lambda x: x.astype(np.float32))
    w_lp = to_lp(rng.uniform(-1, 1, E).astype(np.float32))
    w32 = to_f32(w_lp)
    kw = dict(learning_rate=0.01, wd=0.01)
    if name == "lars":
        kw.update(momentum=0.9, eta=0.01)
    kv = mx.kv.create("device")
    kv.init(0, mx.nd.array(w_lp, mx.gpu(0), dtype=lp))
    kv.set_optimizer(mx.optimizer.create(name, multi_precision=True, **kw))
    out = mx.nd.empty((E,), mx.gpu(0), dtype=lp)
    mean, var = np.zeros(E, np.float32), np.zeros(E, np.float32)
    for t in range(1, 4):
        g_lp = [to_lp(rng.uniform(-1, 1, E).astype(np.float32)) for _ in range(n)]
        kv.pushpull(0, [mx.nd.array(x, mx.gpu(0), dtype=lp) for x in g_lp], out=out)
        gsum = O.sum_device_lp_f32out(g_lp, kind)
        stored = to_f32(to_lp(w32))        # the 16-bit weight the store holds before this step
        if name == "lamb":
            O.lamb_update(w32, gsum, mean, var, 0.01, 0.01, t, norm_mode="f64")
        elif name == "lans":
            O.lans_update(w32, gsum, mean, var, 0.01, 0.01, t, norm_mode="f64", w_norm_src=stored)
        else:
            lr = O.lars_lr(0.01, stored, gsum, 0.01, eta=0.01, norm_mode="f64")
            O.sgd_mom_update(w32, gsum, mean, lr, 0.01, 0.9)
        master = _state(kv, 0, 1)
        np.testing.assert_allclose(master, w32, rtol=RTOL, atol=ATOL, err_msg="%s master step %d" % (name, t))
        # the copy-out is the rounded master, exactly
        got = out.asnumpy(raw=True) if kind == 2 else out.asnumpy().view(np.uint16)
        want = to_lp(master)
        want = want if kind == 2 else want.view(np.uint16)
        assert np.array_equal(got, want)
        w32 = master.copy()               # keep following the device trajectory (no drift accumulation)


def test_skip_nonfinite_leaves_everything_untouched():
    """AMP overflow skip (gluon/trainer.py:445-448) decided on the device: a push whose merged gradient
    holds inf/nan changes neither weight nor state nor update count, for every key of the call."""
    for name in ("lamb", "lans", "lars"):
        rng = _rng(23)
        shapes = [(4099,), (300007,)]
        w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
        kw = dict(learning_rate=0.01, wd=0.01)
        if name == "lars":
            kw.update(momentum=0.9, eta=0.01)
        kv = mx.kv.create("device")
        kv.init([0, 1], [mx.nd.array(w, mx.gpu(0)) for w in w0])
        kv.set_optimizer(mx.optimizer.create(name, skip_nonfinite=True, **kw))
        oopt = O.OracleOptimizer(name, norm_mode="f64", **kw)
        ow = [w.copy() for w in w0]
        outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]

        def step(grads):
            kv.pushpull([0, 1], [[mx.nd.array(g, mx.gpu(0)) for g in gs] for gs in grads], out=outs)

        good = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
        step(good)
        assert kv.overflow() is False
        for k in range(2):
            oopt.update(k, ow[k], O.sum_device(good[k]))
            np.testing.assert_allclose(outs[k].asnumpy(), ow[k], rtol=RTOL, atol=ATOL)
        snap_w = [o.asnumpy().copy() for o in outs]
        snap_s = [_state(kv, k, 2).copy() for k in range(2)]
        # overflow in the SMALL key only: the large key must be skipped as well
        bad = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
        bad[0][1][123] = np.inf
        step(bad)
        for k in range(2):
            assert _bits_equal(outs[k].asnumpy(), snap_w[k]), "%s key %d changed on overflow" % (name, k)
            assert _bits_equal(_state(kv, k, 2), snap_s[k]), "%s state %d changed on overflow" % (name, k)
        assert kv.overflow() is True
        assert kv.overflow() is False            # reading clears it
        # the skipped step does not count: the next good step is step 2 of the oracle
        good2 = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
        step(good2)
        for k in range(2):
            oopt.update(k, ow[k], O.sum_device(good2[k]))
            np.testing.assert_allclose(outs[k].asnumpy(), ow[k], rtol=RTOL, atol=ATOL,
                                       err_msg="%s key %d after a skipped step" % (name, k))
        # NaN produced by the reduction itself (inf + -inf)
        bad2 = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(2)] for s in shapes]
        bad2[1][0][5] = np.inf; bad2[1][1][5] = -np.inf
        snap_w = [o.asnumpy().copy() for o in outs]
        step(bad2)
        assert kv.overflow() is True
        for k in range(2):
            assert _bits_equal(outs[k].asnumpy(), snap_w[k])


def test_lamb_after_gradient_compression():
    """2-bit compressed exchange (comm.h:556-605) feeding the LAMB sequence: the dequantised sum is what
    the optimizer sees, on every consumer GPU."""
    E, n, thr = 10000, 4, 0.5
    rng = _rng(31)
    w0 = rng.uniform(-1, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.01, wd=0.01)
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": "2bit", "threshold": thr})
    kv.init(0, mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.LAMB(**kw))
    oopt = O.OracleOptimizer("lamb", norm_mode="f64", **kw)
    ow = w0.copy()
    residual = [np.zeros(E, np.float32) for _ in range(n)]
    out = mx.nd.empty((E,), mx.gpu(0))
    for step in range(3):
        grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
        kv.pushpull(0, [mx.nd.array(g, mx.gpu(0)) for g in grads], out=out)
        deq = [O.dequantize_2bit(O.quantize_2bit(g, r, thr), E, thr) for g, r in zip(grads, residual)]
        oopt.update(0, ow, O.sum_device(deq))
        np.testing.assert_allclose(out.asnumpy(), ow, rtol=RTOL, atol=ATOL, err_msg="step %d" % step)


def test_host_resident_values_and_outputs():
    """kv.create('local')-style use: gradients and outputs in host memory (staged through the GPU; the
    segment pipeline of the one-pass optimizers cannot be used -- a norm needs the whole key)."""
    E, n = 300007, 2
    rng = _rng(37)
    w0 = rng.uniform(-1, 1, E).astype(np.float32)
    for name, kw in (("lars", dict(learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01)),
                     ("lamb", dict(learning_rate=0.01, wd=0.01))):
        kv = mx.kv.create("local")
        kv.init("w", mx.nd.array(w0))
        kv.set_optimizer(mx.optimizer.create(name, **kw))
        oopt = O.OracleOptimizer(name, norm_mode="f64", **kw)
        ow = w0.copy()
        out = mx.nd.empty((E,), mx.cpu())
        for step in range(2):
            grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
            kv.pushpull("w", [mx.nd.array(g) for g in grads], out=out)
            oopt.update(0, ow, O.sum_cpu([g.copy() for g in grads], 4).reshape(E))
            np.testing.assert_allclose(out.asnumpy(), ow, rtol=RTOL, atol=ATOL, err_msg="%s step %d" % (name, step))


def test_states_round_trip_through_checkpoint(tmp_path):
    """save / load_optimizer_states with LAMB's mean / var (kvstore.py:647-672)."""
    E = 20011
    rng = _rng(29)
    w0 = rng.uniform(-1, 1, E).astype(np.float32)
    kw = dict(learning_rate=0.01, wd=0.01)
    grads = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(4)]

    def run(kv, gs, out):
        for g in gs:
            kv.pushpull(0, mx.nd.array(g, mx.gpu(0)), out=out)

    kv = mx.kv.create("device")
    kv.init(0, mx.nd.array(w0, mx.gpu(0)))
    kv.set_optimizer(mx.optimizer.LAMB(**kw))
    out = mx.nd.empty((E,), mx.gpu(0))
    run(kv, grads[:2], out)
    f = str(tmp_path / "lamb.states")
    kv.save_optimizer_states(f, dump_optimizer=True)      # the optimizer carries the update counts
    mid = out.asnumpy().copy()
    run(kv, grads[2:], out)
    want = out.asnumpy().copy()

    kv2 = mx.kv.create("device")
    kv2.init(0, mx.nd.array(mid, mx.gpu(0)))
    kv2.set_optimizer(mx.optimizer.LAMB(**kw))
    kv2.load_optimizer_states(f)
    out2 = mx.nd.empty((E,), mx.gpu(0))
    run(kv2, grads[2:], out2)
    assert _bits_equal(out2.asnumpy(), want)
