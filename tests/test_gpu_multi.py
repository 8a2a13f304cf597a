"""Multi-GPU parity (skipped below 2 GPUs).

* single process, n GPUs -- the reference's own deployment shape (kv.push(key, [g_gpu0, g_gpu1, ...]));
* one process per GPU (torchrun + NCCL bootstrap, data path = engine kernels over NVLink peer memory).
Everything is compared bit-for-bit with the CPU oracle's CommDevice order.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

import mxnet_b200 as mx
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _devs():
    return list(range(min(mx.num_gpus(), 8)))


@pytest.mark.parametrize("E", [1, 1000, 4099, 65536, 70001, (1 << 20) + 77, 5_000_003])
def test_sp_allreduce_matches_device_order(E):
    """tests/python/gpu/test_device.py:37-60 generalised: push one value per GPU, pull on every GPU."""
    devs = _devs()
    rng = np.random.default_rng(E)
    vals = [rng.uniform(-1, 1, E).astype(np.float32) for _ in devs]
    kv = mx.kv.create("device")
    kv.init(3, mx.nd.zeros((E,), mx.gpu(devs[0])))
    kv.push(3, [mx.nd.array(v, mx.gpu(d)) for v, d in zip(vals, devs)])
    outs = [mx.nd.empty((E,), mx.gpu(d)) for d in devs]
    kv.pull(3, out=outs)
    want = O.sum_device(vals)
    for o in outs:
        assert _bits_equal(o.asnumpy(), want)
    # fused pushpull, then the in-place form (out=None -> values are overwritten with the sum)
    outs2 = [mx.nd.empty((E,), mx.gpu(d)) for d in devs]
    arrs = [mx.nd.array(v, mx.gpu(d)) for v, d in zip(vals, devs)]
    kv.pushpull(3, arrs, out=outs2)
    for o in outs2:
        assert _bits_equal(o.asnumpy(), want)
    kv.pushpull(3, arrs)
    for a in arrs:
        assert _bits_equal(a.asnumpy(), want)


def test_sp_test_device_kat():
    """tests/python/gpu/test_device.py:25-60: shapes 10..100000 and a 7-d shape, result == n_gpus."""
    devs = _devs()
    shapes = [(10,), (100,), (1000,), (10000,), (100000,), (2, 2), (2, 3, 4, 5, 6, 7, 8)]
    keys = list(range(len(shapes)))
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(0)) for s in shapes])
    kv.broadcast([k + 100 for k in keys], [mx.nd.ones(s, mx.gpu(0)) for s in shapes],
                 out=[[mx.nd.empty(s, mx.gpu(d)) for d in devs] for s in shapes])
    vals = [[mx.nd.ones(s, mx.gpu(d)) for d in devs] for s in shapes]
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in devs] for s in shapes]
    kv.pushpull(keys, vals, out=outs)
    for oo in outs:
        for o in oo:
            assert np.all(o.asnumpy() == len(devs))


def test_sp_nccl_store_kat():
    """tests/python/gpu/test_nccl.py:37-52 (skipped upstream unless libmxnet was built with NCCL): kv.create('nccl'),
    string keys, ones pushed from 1 ... n GPUs and pulled back to all of them == the number of GPUs.  The name
    is served by the same peer-memory kernels."""
    shapes = [(10,), (100,), (1000,), (10000,), (100000,), (2, 2), (2, 3, 4, 5, 6, 7, 8)]
    ngpu = min(mx.num_gpus(), 8)
    for key, shape in enumerate(shapes, 1):
        for n_gpus in range(1, ngpu + 1):
            kv = mx.kv.create("nccl")
            assert kv.type == "nccl"
            cur_key = str(key * ngpu + n_gpus)
            kv.init(cur_key, mx.nd.ones(shape, mx.gpu(0)))
            arr_list = [mx.nd.ones(shape, mx.gpu(x)) for x in range(n_gpus)]
            res = [mx.nd.zeros(shape, mx.gpu(x)) for x in range(n_gpus)]
            kv.push(cur_key, arr_list)
            kv.pull(cur_key, res)
            for x in range(n_gpus):
                assert np.sum(np.abs(res[x].asnumpy() - n_gpus)) == 0, (shape, n_gpus, x)


@pytest.mark.parametrize("optname,kw", [
    ("sgd", dict(learning_rate=0.1, momentum=0.9, wd=1e-4, rescale_grad=0.5, clip_gradient=0.7)),
    ("adam", dict(learning_rate=0.01, wd=1e-3)),
])
def test_sp_fused_update_multi_key(optname, kw):
    devs = _devs()
    shapes = [(64,), (1000, 33), (300, 1000), (7,), (1 << 21,)]       # one-shot and two-shot keys
    keys = list(range(len(shapes)))
    rng = np.random.default_rng(17)
    w0 = [rng.uniform(0, 1, s).astype(np.float32) for s in shapes]
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    kv.set_optimizer(mx.optimizer.create(optname, **kw))
    okv = O.OracleKVStore("device")
    okv.init(keys, [w.copy() for w in w0])
    okv.set_optimizer(O.OracleOptimizer(optname, **kw))
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in devs] for s in shapes]
    for step in range(3):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in devs] for s in shapes]
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(d)) for g, d in zip(gs, devs)] for gs in grads], out=outs)
        okv.push(keys, grads)
        for k in keys:
            want = np.empty(shapes[k], np.float32)
            okv.pull(k, want)
            for o in outs[k]:
                assert _bits_equal(o.asnumpy(), want), (optname, step, k)


@pytest.mark.parametrize("optname,kw", [
    ("lamb", dict(learning_rate=0.01, wd=0.01)),
    ("lamb", dict(learning_rate=0.01, wd=0.01, skip_nonfinite=True)),
    ("lans", dict(learning_rate=0.01, wd=0.01, clip_gradient=0.5)),
    ("lars", dict(learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01)),
])
def test_sp_layerwise_adaptive_multi_key(optname, kw):
    """LAMB / LANS / LARS over one-shot and two-shot (sharded: norms added across the GPUs' shards)
    keys.  Tolerance: see tests/test_gpu_norm_opt.py; replicas on different GPUs must be bit-identical."""
    devs = _devs()
    shapes = [(64,), (1000, 33), (300, 1000), (7,), (1 << 21,)]
    keys = list(range(len(shapes)))
    rng = np.random.default_rng(19)
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    kv = mx.kv.create("device")
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    kv.set_optimizer(mx.optimizer.create(optname, **kw))
    oopt = O.OracleOptimizer(optname, norm_mode="f64", **{k: v for k, v in kw.items() if k != "skip_nonfinite"})
    ow = [w.copy() for w in w0]
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in devs] for s in shapes]
    for step in range(3):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in devs] for s in shapes]
        overflow = kw.get("skip_nonfinite") and step == 1
        if overflow:
            grads[3][-1][2] = np.nan
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(d)) for g, d in zip(gs, devs)] for gs in grads], out=outs)
        if kw.get("skip_nonfinite"):
            assert kv.overflow() == bool(overflow)
        for k in keys:
            if not overflow:
                oopt.update(k, ow[k], O.sum_device(grads[k]).reshape(shapes[k]))
            first = outs[k][0].asnumpy()
            np.testing.assert_allclose(first, ow[k], rtol=2e-6, atol=2e-7, err_msg=str((optname, step, k)))
            for o in outs[k][1:]:
                assert _bits_equal(o.asnumpy(), first), (optname, step, k)


def test_sp_save_load_states_sharded(tmp_path):
    """two-shot keys keep optimizer state sharded across the GPUs; save gathers it."""
    devs = _devs()
    E = 1 << 20
    kw = dict(learning_rate=0.01)
    rng = np.random.default_rng(4)
    w0 = rng.uniform(0, 1, E).astype(np.float32)

    def steps(kv, n, seed):
        r = np.random.default_rng(seed)
        outs = [mx.nd.empty((E,), mx.gpu(d)) for d in devs]
        for _ in range(n):
            kv.pushpull(0, [mx.nd.array(r.uniform(-1, 1, E).astype(np.float32), mx.gpu(d)) for d in devs], out=outs)
        return outs[0].asnumpy()

    kv1 = mx.kv.create("device"); kv1.init(0, mx.nd.array(w0, mx.gpu(0))); kv1.set_optimizer(mx.optimizer.Adam(**kw))
    mid = steps(kv1, 2, 1)
    f = str(tmp_path / "s")
    kv1.save_optimizer_states(f, dump_optimizer=True)     # the optimizer carries the update counts
    ref = steps(kv1, 2, 2)
    kv2 = mx.kv.create("device"); kv2.init(0, mx.nd.array(mid, mx.gpu(0))); kv2.set_optimizer(mx.optimizer.Adam(**kw))
    kv2.load_optimizer_states(f)
    got = steps(kv2, 2, 2)
    assert _bits_equal(got, ref)


def test_sp_python_updater_callback():
    """test_kvstore.py:222-274 on GPUs: updater local += recv."""
    devs = _devs()
    shape = (4, 4)
    kv = mx.kv.create("device")
    kv.init(3, mx.nd.zeros(shape, mx.gpu(0)))

    def updater(key, recv, local):
        assert isinstance(key, int)
        local += recv
    kv._set_updater(updater)
    for it in range(1, 4):
        kv.push(3, [mx.nd.ones(shape, mx.gpu(d)) for d in devs])
        outs = [mx.nd.empty(shape, mx.gpu(d)) for d in devs]
        kv.pull(3, out=outs)
        for o in outs:
            assert np.all(o.asnumpy() == it * len(devs))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_mp_one_process_per_gpu(world):
    if mx.num_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("MP_WORKER_OK") == world


@pytest.mark.parametrize("world,local_world", [(4, 2), (8, 4), (2, 1)])
def test_one_process_per_gpu_multi_node_hierarchy(world, local_world):
    """kv.create('dist_device_sync') with the box split into "nodes" of `local_world` GPUs: NVLink peer memory
    inside a node (the engine's own kernels), NCCL all-reduces between the nodes -- the production layout, minus the
    network.  tests/dist_worker.py checks itself against the oracle."""
    if mx.num_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["MXKV_TEST_LOCAL_WORLD"] = str(local_world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("DIST_WORKER_OK") == world
