"""The reference's own KVStore unit tests, re-expressed against the CUDA engine through the C ABI.
Each test cites the reference test it mirrors.  `dev` = 'gpu': every array on GPU 0 (four values on
one device, like the reference's fake multi-context trick); `dev` = 'cpu': host arrays with the
reference's `mx.Context('cpu', i)` contexts (the engine stages them and computes on the GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mxnet_b200 as mx
from oracle import oracle as O

shape = (4, 4)
keys = [5, 7, 11]
str_keys = ["b", "c", "d"]


def ctx_of(dev, i=0):
    return mx.gpu(0) if dev == "gpu" else mx.Context("cpu", i)


def init_kv(dev, name="device"):
    kv = mx.kv.create(name)
    kv.init(3, mx.nd.zeros(shape, ctx_of(dev)))
    kv.init(keys, [mx.nd.zeros(shape, ctx_of(dev))] * len(keys))
    return kv


def init_kv_with_str(dev, name="device"):
    kv = mx.kv.create(name)
    kv.init("a", mx.nd.zeros(shape, ctx_of(dev)))
    kv.init(str_keys, [mx.nd.zeros(shape, ctx_of(dev))] * len(keys))
    return kv


def check_diff_to_scalar(A, x):
    assert np.sum(np.abs(A.asnumpy() - x)) == 0, (A.asnumpy(), x)


@pytest.mark.parametrize("dev", ["gpu", "cpu"])
@pytest.mark.parametrize("name", ["device", "local"])
def test_single_kv_pair_init_pull_list(dev, name):
    # tests/python/unittest/test_kvstore.py:55-66
    for kv, key in ((init_kv(dev, name), 3), (init_kv_with_str(dev, name), "a")):
        kv.push(key, mx.nd.ones(shape, ctx_of(dev)))
        val = mx.nd.empty(shape, ctx_of(dev))
        kv.pull(key, out=val)
        check_diff_to_scalar(val, 1)
    # :96-105
    for key in (3, "a"):
        kv = mx.kv.create(name)
        kv.init(key, mx.nd.ones(shape, ctx_of(dev)) * 4)
        a = mx.nd.zeros(shape, ctx_of(dev))
        kv.pull(key, out=a)
        check_diff_to_scalar(a, 4)
    # :107-121
    kv = mx.kv.create(name)
    a = mx.nd.ones(shape, ctx_of(dev))
    b = mx.nd.zeros(shape, ctx_of(dev))
    kv.init("1", mx.nd.zeros(shape, ctx_of(dev)))
    kv.push("1", [a, a, a, a])
    kv.pull("1", b)
    check_diff_to_scalar(b, 4)
    kv.init("2", mx.nd.zeros(shape, ctx_of(dev)))
    kv.pull("2", b)
    check_diff_to_scalar(b, 0)
    # :123-136
    for kv, kl in ((init_kv(dev, name), keys), (init_kv_with_str(dev, name), str_keys)):
        kv.push(kl, [mx.nd.ones(shape, ctx_of(dev)) * 4] * len(kl))
        val = [mx.nd.empty(shape, ctx_of(dev))] * len(kl)
        kv.pull(kl, out=val)
        for v in val:
            check_diff_to_scalar(v, 4)


@pytest.mark.parametrize("dev", ["gpu", "cpu"])
def test_updater(dev):
    # test_kvstore.py:222-274
    def updater(key, recv, local):
        assert isinstance(key, int)
        local += recv

    def str_updater(key, recv, local):
        assert isinstance(key, str)
        local += recv

    def check_updater(kv, key, key_list):
        num_devs = 4
        devs = [ctx_of(dev, i) for i in range(num_devs)]
        vals = [mx.nd.ones(shape, d) for d in devs]
        outs = [mx.nd.empty(shape, d) for d in devs]
        kv.push(key, vals)
        kv.pull(key, out=outs)
        for out in outs:
            check_diff_to_scalar(out, num_devs)
        vals = [[mx.nd.ones(shape, d) for d in devs]] * len(key_list)
        outs = [[mx.nd.empty(shape, d) for d in devs]] * len(key_list)
        num_push = 4
        for _ in range(num_push):
            kv.push(key_list, vals)
        kv.pull(key_list, out=outs)
        for out in outs:
            for o in out:
                check_diff_to_scalar(o, num_devs * num_push)

    kv = init_kv(dev)
    kv._set_updater(updater)
    check_updater(kv, 3, keys)
    skv = init_kv_with_str(dev)
    skv._set_updater(str_updater)
    check_updater(skv, "a", str_keys)


def test_get_type_and_invalid_ops():
    # test_kvstore.py:276-279, :281-339
    assert mx.kv.create("local_allreduce_cpu").type == "local_allreduce_cpu"
    int_kv, str_kv = init_kv("gpu"), init_kv_with_str("gpu")
    dns = mx.nd.ones(shape, mx.gpu(0)) * 2
    for kv, bad in ((int_kv, "a"), (str_kv, 3)):
        for fn in (lambda: kv.init(bad, dns), lambda: kv.push(bad, dns), lambda: kv.pull(bad, dns)):
            with pytest.raises(mx.MXNetError):
                fn()


@pytest.mark.parametrize("dev", ["gpu", "cpu"])
def test_broadcast_and_pushpull_custom_api(dev):
    # tests/python/unittest/test_kvstore_custom.py:37-152 ('device' rows)
    for key in (3, "a"):
        kv = mx.kv.create("device")
        ones = mx.nd.ones(shape, ctx_of(dev))
        out = mx.nd.empty(shape, ctx_of(dev))
        kv.broadcast(key, ones, out)
        check_diff_to_scalar(out, 1)
        out_list = [mx.nd.empty(shape, ctx_of(dev))] * 3
        kv.broadcast(key + key, ones, out_list)
        for o in out_list:
            check_diff_to_scalar(o, 1)
    for kl in (keys, str_keys):
        kv = mx.kv.create("device")
        ones = [mx.nd.ones(shape, ctx_of(dev))] * len(kl)
        out = [mx.nd.empty(shape, ctx_of(dev))] * len(kl)
        kv.broadcast(kl, ones, out)
        for o in out:
            check_diff_to_scalar(o, 1)
        out_list = [[mx.nd.empty(shape, ctx_of(dev))] * 2 for _ in kl]
        kv.broadcast([k + k for k in kl], ones, out_list)
        for o in out_list:
            for oo in o:
                check_diff_to_scalar(oo, 1)
    for key, key_list in ((3, keys), ("a", str_keys)):
        kv = mx.kv.create("device")
        kv.broadcast(key, mx.nd.zeros(shape, ctx_of(dev)), out=mx.nd.empty(shape, ctx_of(dev)))
        num_devs = 4
        devs = [ctx_of(dev, i) for i in range(num_devs)]
        vals = [mx.nd.ones(shape, d) for d in devs]
        outs = [mx.nd.empty(shape, d) for d in devs]
        kv.pushpull(key, vals, out=outs)
        for out in outs:
            check_diff_to_scalar(out, num_devs)
        kv.pushpull(key, vals)                         # in place
        for val in vals:
            check_diff_to_scalar(val, num_devs)
        kv.broadcast(key_list, [mx.nd.zeros(shape, ctx_of(dev))] * len(key_list),
                     out=[mx.nd.empty(shape, ctx_of(dev))] * len(key_list))
        vals = [[mx.nd.ones(shape, d) * 2.0 for d in devs]] * len(key_list)
        outs = [[mx.nd.empty(shape, d) for d in devs]] * len(key_list)
        kv.pushpull(key_list, vals, out=outs)
        for out in outs:
            for o in out:
                check_diff_to_scalar(o, num_devs * 2.0)
        kv.pushpull(key_list, vals)
        for val in vals:
            for v in val:
                check_diff_to_scalar(v, num_devs * 2.0)


def test_nightly_random_accuracy():
    # tests/nightly/test_kvstore.py:297-343: 4 workers, 3 repeats, rel-L1 < 1e-6 against float64 numpy
    rng = np.random.default_rng(0)
    shapes = [(4, 4), (100, 100), (2000, 2000)]
    kv = mx.kv.create("device")
    kv.set_optimizer(mx.optimizer.create("test", rescale_grad=0.1))
    ks = [3, 5, 7]
    for k, s in zip(ks, shapes):
        kv.init(k, mx.nd.zeros(s, mx.gpu(0)))
    res = [np.zeros(s) for s in shapes]
    lr = 0.01
    for _ in range(3):
        for j, s in enumerate(shapes):
            data = [rng.uniform(-1, 1, s).astype(np.float32) for _ in range(4)]
            kv.push(ks[j], [mx.nd.array(d, mx.gpu(0)) for d in data])
            res[j] = res[j] - lr * 0.1 * sum(d.astype(np.float64) for d in data)
        for j, s in enumerate(shapes):
            out = mx.nd.zeros(s, mx.gpu(0))
            kv.pull(ks[j], out=out)
            err = np.sum(np.abs(out.asnumpy() - res[j])) / np.sum(np.abs(res[j]))
            assert err < 1e-6, (err, s)


@pytest.mark.parametrize("stype", ["default", "row_sparse"])
def test_aggregator(stype):
    # tests/python/unittest/test_kvstore.py:139-172: values on four cpu contexts, single key then a key
    # list, int and str keys; row_sparse values are pushed to keys that were initialised dense
    def check_aggregator(kv, key, key_list):
        num_devs = 4
        devs = [mx.Context("cpu", i) for i in range(num_devs)]
        vals = [mx.nd.ones(shape, d).tostype(stype) for d in devs]
        outs = [mx.nd.empty(shape, d) for d in devs]
        kv.push(key, vals)
        kv.pull(key, out=outs)
        for out in outs:
            check_diff_to_scalar(out, num_devs)
        vals = [[(mx.nd.ones(shape, d) * 2.0).tostype(stype) for d in devs]] * len(key_list)
        outs = [[mx.nd.empty(shape, d) for d in devs]] * len(key_list)
        kv.push(key_list, vals)
        kv.pull(key_list, out=outs)
        for out in outs:
            for o in out:
                check_diff_to_scalar(o, num_devs * 2.0)

    check_aggregator(init_kv("cpu"), 3, keys)
    check_aggregator(init_kv_with_str("cpu"), "a", str_keys)


@pytest.mark.parametrize("sparse_pull", [False, True])
@pytest.mark.parametrize("dev", ["cpu", "gpu"])
def test_sparse_aggregator(sparse_pull, dev):
    # tests/python/unittest/test_kvstore.py:174-220: row_sparse keys, random row_sparse values on four contexts,
    # pushed and then read back INTO THE SAME ARRAYS, either with row_sparse_pull of every row or with
    # pull(ignore_sparse=False); single key, then the key list with one shared list of values
    rng = np.random.default_rng(11 + int(sparse_pull))

    def rand_rsp(ctx):
        dense = rng.normal(size=shape).astype(np.float32)
        dense[rng.random(shape[0]) < 0.5] = 0                     # rand_ndarray: random density
        return mx.nd.array(dense, ctx).tostype("row_sparse")

    kv = mx.kv.create("device")
    kv.init("a", mx.nd.zeros(shape, stype="row_sparse"))
    kv.init(str_keys, [mx.nd.zeros(shape, stype="row_sparse")] * len(keys))
    num_devs = 4
    devs = [ctx_of(dev, i) for i in range(num_devs)]
    all_rows = mx.nd.array(np.arange(shape[0]), dtype=np.float32)

    vals = [rand_rsp(d) for d in devs]
    expected_sum = np.zeros(shape)
    for v in vals:
        expected_sum += v.todense_numpy()
    kv.push("a", vals)
    if sparse_pull:
        kv.row_sparse_pull("a", out=vals, row_ids=[all_rows] * len(vals))
    else:
        kv.pull("a", out=vals, ignore_sparse=False)
    result_sum = np.zeros(shape)
    for v in vals:
        result_sum += v.todense_numpy()
    np.testing.assert_allclose(result_sum, expected_sum * num_devs, rtol=1e-5, atol=1e-6)

    vals = [[rand_rsp(d) for d in devs]] * len(keys)
    expected_sum = np.zeros(shape)
    for v in vals[0]:
        expected_sum += v.todense_numpy()
    kv.push(str_keys, vals)
    if sparse_pull:
        kv.row_sparse_pull(str_keys, out=vals, row_ids=[[all_rows] * num_devs] * len(vals))
    else:
        kv.pull(str_keys, out=vals, ignore_sparse=False)
    for vv in vals:
        result_sum = np.zeros(shape)
        for v in vv:
            result_sum += v.todense_numpy()
        np.testing.assert_allclose(result_sum, expected_sum * num_devs, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kv_type", ["local", "device"])
@pytest.mark.parametrize("push_on_gpu", [False, True])
def test_rsp_push_pull(kv_type, push_on_gpu):
    # tests/python/gpu/test_kvstore_gpu.py:48-110 on one GPU: row_sparse key of ones, two row_sparse
    # pushes of ones (stored value becomes 2), then row_sparse_pull with random -- repeated, unsorted,
    # FLOAT32 -- row ids into outputs on the GPU and on cpu contexts, one id array per output or a
    # shared one, and the dense pull of the whole value
    rshape = (20, 6)
    num_rows = rshape[0]
    kv = mx.kv.create(kv_type)
    kv.init("a", mx.nd.zeros(rshape, stype="row_sparse"))
    kv.init("e", mx.nd.ones(rshape).tostype("row_sparse"))
    push_ctxs = [mx.gpu(0) if push_on_gpu else mx.Context("cpu", i) for i in range(2)]
    kv.push("e", [mx.nd.ones(rshape, c).tostype("row_sparse") for c in push_ctxs])
    rng = np.random.default_rng(3)

    def check_rsp_pull(ctxs, is_same_rowid=False):
        count = len(ctxs)
        all_row_ids = np.arange(num_rows)
        vals = [mx.nd.zeros(rshape, c, stype="row_sparse") for c in ctxs]
        if is_same_rowid:
            row_id = rng.integers(0, num_rows, num_rows)
            row_ids = [mx.nd.array(row_id.astype(np.float32), dtype=np.float32)] * count
        else:
            row_ids = [mx.nd.array(rng.integers(0, num_rows, num_rows).astype(np.float32), dtype=np.float32)
                       for _ in range(count)]
        row_ids_to_pull = row_ids[0] if (len(row_ids) == 1 or is_same_rowid) else row_ids
        vals_to_pull = vals[0] if len(vals) == 1 else vals
        kv.row_sparse_pull("e", out=vals_to_pull, row_ids=row_ids_to_pull)
        for val, row_id in zip(vals, row_ids):
            retained = val.todense_numpy()
            excluded = np.setdiff1d(all_row_ids, row_id.asnumpy())
            for row in range(num_rows):
                want = 0.0 if row in excluded else 2.0
                assert np.all(retained[row] == want), (row, retained[row], want)
        kv.pull("e", out=vals_to_pull, ignore_sparse=False)
        for val in vals:
            assert np.all(val.todense_numpy() == 2.0)

    check_rsp_pull([mx.gpu(0)])
    check_rsp_pull([mx.cpu(0)])
    check_rsp_pull([mx.gpu(0) for _ in range(4)])
    check_rsp_pull([mx.gpu(0) for _ in range(4)], is_same_rowid=True)
    check_rsp_pull([mx.Context("cpu", i) for i in range(4)])
    check_rsp_pull([mx.Context("cpu", i) for i in range(4)], is_same_rowid=True)


def test_row_sparse_pull_single_device_and_large_rowid():
    # tests/python/gpu/test_kvstore_gpu.py:112-125
    rng = np.random.default_rng(4)
    dense = rng.normal(size=(4, 4)).astype(np.float32)
    grad = mx.nd.array(dense, mx.gpu(0)).tostype("row_sparse")
    kv = mx.kv.create("device")
    kv.init(0, grad)
    idx = grad.indices
    kv.push(0, grad)
    kv.row_sparse_pull(0, out=grad, row_ids=idx)
    assert np.array_equal(grad.todense_numpy(), dense)
    # :127-136: 793470 rows of one element, every row id requested
    num_rows = 793470
    val = mx.nd.row_sparse_array(np.ones((num_rows, 1), np.float32), ctx=mx.gpu(0))
    kv = mx.kv.create("device")
    kv.init("a", val)
    out = mx.nd.zeros((num_rows, 1), mx.gpu(0), stype="row_sparse")
    kv.push("a", val)
    kv.row_sparse_pull("a", out=out, row_ids=mx.nd.array(np.arange(num_rows, dtype=np.int64), mx.gpu(0), dtype=np.int64))
    assert out.indices.shape[0] == num_rows
    assert np.all(out.data.asnumpy() == 1.0)


def test_user_defined_python_optimizer_on_the_store():
    # kvstore.py:559-606: an optimizer without a fused kernel runs through the updater callback, on the merged
    # value, once per pushed key; written against the reference's Optimizer protocol (list-valued step that
    # counts the update itself)
    @mx.optimizer.register
    class HalfStep(mx.optimizer.Optimizer):
        def create_state(self, index, weight):
            return mx.nd.zeros(weight.shape, weight.context)

        def step(self, indices, weights, grads, states):
            self._update_count(indices)
            for i, w, g, s, lr in zip(indices, weights, grads, states, self._get_lrs(indices)):
                s[:] = s.asnumpy() + 1
                w[:] = w.asnumpy() - lr * self.rescale_grad * g.asnumpy() / s.asnumpy()

    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.ones(shape, mx.gpu(0))] * len(keys))
    opt = mx.optimizer.create("halfstep", learning_rate=0.5, rescale_grad=0.25)
    kv.set_optimizer(opt)
    want = np.ones(shape, np.float32)
    for step in (1, 2, 3):
        kv.push(keys, [[mx.nd.ones(shape, mx.gpu(0)) * 2.0 for _ in range(4)]] * len(keys))
        want = want - np.float32(0.5 * 0.25) * np.float32(8.0) / np.float32(step)
        outs = [mx.nd.empty(shape, mx.gpu(0)) for _ in keys]
        kv.pull(keys, out=outs)
        for o in outs:
            np.testing.assert_allclose(o.asnumpy(), want, rtol=1e-6)
    assert opt._index_update_count == {k: 3 for k in keys} and opt.num_update == 3


@pytest.mark.parametrize("kv_store,optimizer", [("device", None), ("device", "sgd"), ("local", None), ("local", "sgd")])
def test_bandwidth_tool_results(kv_store, optimizer):
    # tools/bandwidth/test_measure.py:30-44 over tools/bandwidth/measure.py:76-152: the ResNet-50 key set, one
    # gradient per GPU, per-key push(i, g, priority=i) then pull(i, w, priority=i) for two batches; the relative L1
    # error against numpy-summed gradients (and a host-side SGD updater when an optimizer is set) stays below 1e-4
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    shapes = bench.keyset("resnet50")
    devs = [mx.gpu(i) for i in range(max(1, min(mx.num_gpus(), 8)))]
    rng = np.random.default_rng(50)
    kv = mx.kv.create(kv_store)
    oopt = None
    if optimizer is not None:
        kv.set_optimizer(mx.optimizer.create(optimizer))
        oopt = O.OracleOptimizer(optimizer)
    for i, s in enumerate(shapes):
        kv.init(i, mx.nd.zeros(s))
    grads_val = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in devs] for s in shapes]
    grads = [[mx.nd.array(g, d) for g, d in zip(gs, devs)] for gs in grads_val]
    weights = [[mx.nd.zeros(s, d) for d in devs] for s in shapes]
    cpu_grads = [np.sum(np.stack(gs).astype(np.float64), axis=0).astype(np.float32) for gs in grads_val]
    cpu_weights = [np.zeros(s, np.float32) for s in shapes]
    for _ in range(2):
        for i, g in enumerate(grads):
            kv.push(i, g, i)
        for i, w in enumerate(weights):
            kv.pull(i, w, i)
        if oopt is None:
            want = cpu_grads
        else:
            for i in range(len(shapes)):
                oopt.update(i, cpu_weights[i], cpu_grads[i])
            want = cpu_weights
        num = sum(np.sum(np.abs(a.asnumpy() - b)) for w, b in zip(weights, want) for a in w)
        den = sum(np.sum(np.abs(b)) for b in want)
        assert num / den < 1e-4, (kv_store, optimizer, num / den)
