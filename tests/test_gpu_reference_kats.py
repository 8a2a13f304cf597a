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


@pytest.mark.parametrize("usetree", [None, "1"])
@pytest.mark.parametrize("kv_type", ["local", "device"])
@pytest.mark.parametrize("push_on_gpu", [False, True])
def test_rsp_push_pull(kv_type, push_on_gpu, usetree, monkeypatch):
    # tests/python/gpu/test_kvstore_gpu.py:48-110 on one GPU: row_sparse key of ones, two row_sparse
    # pushes of ones (stored value becomes 2), then row_sparse_pull with random -- repeated, unsorted,
    # FLOAT32 -- row ids into outputs on the GPU and on cpu contexts, one id array per output or a
    # shared one, and the dense pull of the whole value; with and without MXNET_KVSTORE_USETREE like the
    # reference's loop (:100-109; row_sparse keys take CommDevice's own reduce under the tree, comm_tree.h:246-249)
    if usetree is not None:
        monkeypatch.setenv("MXNET_KVSTORE_USETREE", usetree)
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
