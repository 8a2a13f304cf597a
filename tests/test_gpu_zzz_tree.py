"""MXNET_KVSTORE_USETREE=1 (reference: CommDeviceTree, src/kvstore/comm_tree.h:50-325): a `device` store adds the
values of a key pairwise up the binary trees the reference's solver derives from the GPUs' link matrix, slice by
slice above MXNET_KVSTORE_TREE_ARRAY_BOUND.  Every result is compared bit for bit with the oracle's level-by-level
restatement of CommDeviceTree::ReduceInner / Reduce (oracle.tree_reduce / sum_tree) over the trees of THIS machine's
link matrix; the trees themselves are pinned to the reference's compiled solver in tests/test_topology.py.

Needs three GPUs or more: two values add the same way in any order, and the engine does not switch kernels for them.
(File name: written against the simulated runtime after the last hardware session of round 2; sorts after the
hardware-validated files so that `pytest -x` reaches those first.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

import mxnet_b200 as mx
from oracle import oracle as O

T = mx.topology

# the 8-GPU NVLink hybrid cube mesh the reference's solver was written for (gpu_topology.h:181-190), as weights
P3_16XLARGE = np.array([[0, 2, 2, 3, 3, 0, 0, 0], [2, 0, 3, 2, 0, 3, 0, 0], [2, 3, 0, 3, 0, 0, 2, 0],
                        [3, 2, 3, 0, 0, 0, 0, 2], [3, 0, 0, 0, 0, 2, 2, 3], [0, 3, 0, 0, 2, 0, 3, 2],
                        [0, 0, 2, 0, 2, 3, 0, 3], [0, 0, 0, 2, 3, 2, 3, 0]], np.float32)


def _bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def _ngpu():
    return min(mx.num_gpus(), 8)


def _need(n):
    if _ngpu() < max(n, 3):
        pytest.skip("needs %d GPUs" % max(n, 3))


def _tree_store(monkeypatch, bound=None, links=None, backtrack=None, kind="device"):
    monkeypatch.setenv("MXNET_KVSTORE_USETREE", "1")
    if bound is not None:
        monkeypatch.setenv("MXNET_KVSTORE_TREE_ARRAY_BOUND", str(bound))
    if backtrack is not None:
        monkeypatch.setenv("MXNET_KVSTORE_TREE_BACKTRACK", str(backtrack))
    if links is not None:
        monkeypatch.setenv("MXKV_B200_TREE_LINKS", ",".join("%g" % v for v in np.asarray(links).ravel()))
    return mx.kv.create(kind)


def _trees(n, links=None, backtrack=False, bound=10000000):
    W = np.asarray(links, np.float32) if links is not None else T.query_links(list(range(n)))
    topo, scan, depth = T.compute_trees(W, 0.7, backtrack)
    return dict(topo=topo, scan=scan, depth=depth, bound=bound)


def _tree_sum(vals, tree, add=None):
    return O.sum_tree(vals, tree["topo"], tree["scan"], tree["depth"], tree["bound"], add)


SHAPES = [(10,), (1000,), (4099,), (37, 13), (70001,), (64, 33), ((1 << 20) + 77,), (2, 3, 4, 5, 6, 7, 8)]


@pytest.mark.parametrize("bound", [10000000, 100])
@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_tree_allreduce_is_the_reference_tree_sum(monkeypatch, n, bound):
    """push one value per GPU, pull on every GPU (tests/python/gpu/test_device.py:37-60 with random data): whole keys
    go up tree 0, keys above the bound are summed slice by slice up the tree rooted at the slice's GPU -- including
    slices that do not begin on a 16-byte boundary ((37, 13), (64, 33)) and keys the engine shards over the GPUs
    (two-shot, 256 KB and more)."""
    _need(n)
    tree = _trees(n, bound=bound)
    kv = _tree_store(monkeypatch, bound=bound)
    rng = np.random.default_rng(100 * n + (bound < 1000))
    keys = list(range(len(SHAPES)))
    kv.init(keys, [mx.nd.zeros(s, mx.gpu(0)) for s in SHAPES])
    vals = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(n)] for s in SHAPES]
    before = mx.kv.launch_count("tree")
    kv.push(keys, [[mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vs)] for vs in vals])
    assert mx.kv.launch_count("tree") > before, "the tree kernel did not run"
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in range(n)] for s in SHAPES]
    kv.pull(keys, out=outs)
    differs = 0
    for k, s in enumerate(SHAPES):
        want = _tree_sum(vals[k], tree)
        differs += int(not _bits_equal(want, O.sum_device(vals[k])))
        for o in outs[k]:
            assert _bits_equal(o.asnumpy(), want), (s, n, bound)
    assert differs > 0, "every key happened to round like the plain order: the test proves nothing"
    # fused pushpull into fresh outputs and in place
    outs2 = [[mx.nd.empty(s, mx.gpu(d)) for d in range(n)] for s in SHAPES]
    arrs = [[mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vs)] for vs in vals]
    kv.pushpull(keys, arrs, out=outs2)
    kv.pushpull(keys, arrs)
    for k in keys:
        want = _tree_sum(vals[k], tree)
        for o, a in zip(outs2[k], arrs[k]):
            assert _bits_equal(o.asnumpy(), want) and _bits_equal(a.asnumpy(), want)


@pytest.mark.parametrize("bound", [None, 1])
def test_reference_kat_device_pushpull_with_usetree(monkeypatch, bound):
    """tests/python/gpu/test_device.py:39-60 as the reference runs it with MXNET_KVSTORE_USETREE=1 and
    MXNET_KVSTORE_TREE_ARRAY_BOUND unset / 1: ones pushed from 1 ... n GPUs, pulled back everywhere == n."""
    _need(3)
    shapes = [(10,), (100,), (1000,), (10000,), (100000,), (2, 2), (2, 3, 4, 5, 6, 7, 8)]
    for key, shape in enumerate(shapes, 1):
        for n_gpus in range(1, _ngpu() + 1):
            kv = _tree_store(monkeypatch, bound=bound)
            cur_key = str(key * 8 + n_gpus)
            kv.init(cur_key, mx.nd.ones(shape, mx.gpu(0)))
            kv.push(cur_key, [mx.nd.ones(shape, mx.gpu(x)) for x in range(n_gpus)])
            res = [mx.nd.zeros(shape, mx.gpu(x)) for x in range(n_gpus)]
            kv.pull(cur_key, res)
            for r in res:
                assert np.sum(np.abs(r.asnumpy() - n_gpus)) == 0, (shape, n_gpus)


@pytest.mark.parametrize("optname,kw", [
    ("sgd", dict(learning_rate=0.1, momentum=0.9, wd=1e-4, rescale_grad=0.5, clip_gradient=0.7)),
    ("adam", dict(learning_rate=0.01, wd=1e-3)),
])
@pytest.mark.parametrize("n", [3, 4, 8])
def test_tree_sum_feeds_the_fused_optimizer(monkeypatch, n, optname, kw):
    """update_on_kvstore under MXNET_KVSTORE_USETREE: the optimizer sees the tree-order sum (small keys whole, big
    keys by slices), state sharded or not -- three steps against the oracle store in tree mode."""
    _need(n)
    bound = 5000
    tree = _trees(n, bound=bound)
    shapes = [(1000,), (64, 33), (70001,), (300, 257)]
    keys = list(range(len(shapes)))
    rng = np.random.default_rng(n)
    w0 = [rng.uniform(0, 1, s).astype(np.float32) for s in shapes]
    kv = _tree_store(monkeypatch, bound=bound)
    kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
    kv.set_optimizer(getattr(mx.optimizer, "SGD" if optname == "sgd" else "Adam")(**kw))
    okv = O.OracleKVStore("device", tree=tree)
    okv.init(keys, [w.copy() for w in w0])
    okv.set_optimizer(O.OracleOptimizer(optname, **kw))
    outs = [[mx.nd.empty(s, mx.gpu(d)) for d in range(n)] for s in shapes]
    for _ in range(3):
        grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(n)] for s in shapes]
        kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(d)) for d, g in enumerate(gs)] for gs in grads], out=outs)
        okv.push(keys, grads)
    for k, s in enumerate(shapes):
        want = np.empty(s, np.float32)
        okv.pull(k, want)
        for o in outs[k]:
            assert _bits_equal(o.asnumpy(), want), (s, n, optname)


def test_tree_sum_float16_and_float64(monkeypatch):
    """float16 keys round after every pairwise add when no optimizer follows (half_t arithmetic of ElementwiseSum),
    float64 keys add in double; the integer dtypes are associative and stay on the plain kernel."""
    n = min(_ngpu(), 8)
    _need(3)
    tree = _trees(n, bound=2000)
    kv = _tree_store(monkeypatch, bound=2000)
    rng = np.random.default_rng(7)
    shape = (129, 31)
    v16 = [rng.uniform(-1, 1, shape).astype(np.float16) for _ in range(n)]
    v64 = [rng.uniform(-1, 1, shape) for _ in range(n)]
    v32i = [rng.integers(-1000, 1000, shape).astype(np.int32) for _ in range(n)]
    kv.init([0, 1, 2], [mx.nd.zeros(shape, mx.gpu(0), dtype=d) for d in ("float16", "float64", "int32")])
    kv.push([0, 1, 2], [[mx.nd.array(v, mx.gpu(d), dtype=v.dtype) for d, v in enumerate(vs)] for vs in (v16, v64, v32i)])
    outs = [mx.nd.empty(shape, mx.gpu(n - 1), dtype=d) for d in ("float16", "float64", "int32")]
    kv.pull([0, 1, 2], out=outs)
    half_add = lambda a, b: (a.astype(np.float32) + b.astype(np.float32)).astype(np.float16)  # noqa: E731
    assert _bits_equal(outs[0].asnumpy(), _tree_sum(v16, tree, half_add))
    assert _bits_equal(outs[1].asnumpy(), _tree_sum(v64, tree))
    assert np.array_equal(outs[2].asnumpy(), sum(v32i))


def test_float16_multi_precision_under_the_tree(monkeypatch):
    """float16 gradients, float32 master weights: the tree's partial sums are float32 (as the fused path keeps them
    everywhere), rounded once when the new weight is written."""
    n = min(_ngpu(), 8)
    _need(3)
    tree = _trees(n, bound=1000)
    shape = (70, 65)
    rng = np.random.default_rng(11)
    w0 = rng.uniform(0, 1, shape).astype(np.float16)
    kw = dict(learning_rate=0.05, momentum=0.9, wd=1e-4, multi_precision=True)
    kv = _tree_store(monkeypatch, bound=1000)
    kv.init(0, mx.nd.array(w0, mx.gpu(0), dtype=np.float16))
    kv.set_optimizer(mx.optimizer.SGD(**kw))
    w32 = w0.astype(np.float32)
    mom = np.zeros(shape, np.float32)
    oopt = O.OracleOptimizer("sgd", **{k: v for k, v in kw.items() if k != "multi_precision"})
    out = mx.nd.empty(shape, mx.gpu(0), dtype=np.float16)
    for _ in range(2):
        grads = [rng.uniform(-1, 1, shape).astype(np.float16) for _ in range(n)]
        kv.pushpull(0, [mx.nd.array(g, mx.gpu(d), dtype=np.float16) for d, g in enumerate(grads)], out=out)
        g32 = _tree_sum([g.astype(np.float32) for g in grads], tree)
        oopt.states[0] = mom
        oopt.update(0, w32, g32)
        mom = oopt.states[0]
    assert _bits_equal(out.asnumpy(), w32.astype(np.float16))


@pytest.mark.parametrize("backtrack", [0, 1])
def test_trees_follow_the_link_matrix(monkeypatch, backtrack):
    """a non-uniform machine (the NVLink hybrid cube mesh of gpu_topology.h:181-190, injected through
    MXKV_B200_TREE_LINKS) gives other trees, built by Kernighan-Lin or by the exhaustive search
    (MXNET_KVSTORE_TREE_BACKTRACK): the sums follow them."""
    _need(8)
    tree = _trees(8, links=P3_16XLARGE, backtrack=bool(backtrack), bound=3000)
    uniform = _trees(8, bound=3000)
    assert not np.array_equal(tree["topo"], uniform["topo"])
    kv = _tree_store(monkeypatch, bound=3000, links=P3_16XLARGE, backtrack=backtrack)
    rng = np.random.default_rng(5)
    shapes = [(999,), (160, 40)]
    kv.init([0, 1], [mx.nd.zeros(s, mx.gpu(0)) for s in shapes])
    vals = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(8)] for s in shapes]
    outs = [mx.nd.empty(s, mx.gpu(3)) for s in shapes]
    kv.pushpull([0, 1], [[mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vs)] for vs in vals], out=outs)
    for k in (0, 1):
        assert _bits_equal(outs[k].asnumpy(), _tree_sum(vals[k], tree))


def test_what_the_tree_mode_leaves_alone(monkeypatch):
    """two GPUs (any order gives the same bits), one GPU, a `local` store: no tree kernel; a store without the
    variable: plain order; LAMB under the tree: refused."""
    _need(3)
    rng = np.random.default_rng(3)
    vals = [rng.uniform(-1, 1, 5000).astype(np.float32) for _ in range(3)]
    kv = _tree_store(monkeypatch)
    kv.init(0, mx.nd.zeros((5000,), mx.gpu(0)))
    before = mx.kv.launch_count("tree")
    out = mx.nd.empty((5000,), mx.gpu(1))
    kv.pushpull(0, [mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vals[:2])], out=out)
    assert mx.kv.launch_count("tree") == before
    assert _bits_equal(out.asnumpy(), O.sum_device(vals[:2]))
    kvl = _tree_store(monkeypatch, kind="local")
    kvl.init(0, mx.nd.zeros((5000,), mx.gpu(0)))
    kvl.pushpull(0, [mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vals)], out=out)
    assert mx.kv.launch_count("tree") == before
    assert _bits_equal(out.asnumpy(), O.sum_cpu(vals))
    monkeypatch.delenv("MXNET_KVSTORE_USETREE")
    kvp = mx.kv.create("device")
    kvp.init(0, mx.nd.zeros((5000,), mx.gpu(0)))
    kvp.pushpull(0, [mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vals)], out=out)
    assert mx.kv.launch_count("tree") == before
    assert _bits_equal(out.asnumpy(), O.sum_device(vals))
    kvt = _tree_store(monkeypatch)
    kvt.init(0, mx.nd.zeros((5000,), mx.gpu(0)))
    kvt.set_optimizer(mx.optimizer.LAMB(learning_rate=0.01))
    with pytest.raises(mx.MXNetError, match="USETREE"):
        kvt.push(0, [mx.nd.array(v, mx.gpu(d)) for d, v in enumerate(vals)])


_EXTRA = int(os.environ.get("MXKV_FUZZ_SEEDS", "0"))     # more random walks on demand (soak runs on the simulator)


@pytest.mark.parametrize("seed", list(range(12 + _EXTRA)))
def test_randomized_call_sequences_under_the_tree(monkeypatch, seed):
    """The random walk of test_gpu_y_placement.py::test_randomized_call_sequences with the tree on: keys of one-shot and
    two-shot size, some of them sliced, pushed from random subsets of the GPUs -- three or more distinct GPUs add in the
    order of THAT subset's trees, fewer (or a value from the host) in the plain order --, pushes and fused pushpulls
    interleaved with pulls to random devices, a fused optimizer switched on part-way (its state re-laid out when the
    subset changes); compared with the oracle after every call."""
    _need(3)
    ngpu = _ngpu()
    bound = 3000
    trees = {n: _trees(n, bound=bound) for n in range(3, ngpu + 1)}
    rng = np.random.default_rng(7000 + seed)
    all_shapes = [(5,), (64, 10), (4099,), (70001,), (1200, 251), (37, 13)]
    shapes = [all_shapes[i] for i in rng.choice(len(all_shapes), size=3, replace=False)]
    keys = list(range(len(shapes)))
    w0 = [rng.uniform(-1, 1, s).astype(np.float32) for s in shapes]
    kv = _tree_store(monkeypatch, bound=bound)
    okv = O.OracleKVStore("device")
    kv.init(keys, [mx.nd.array(w, mx.gpu(int(rng.integers(ngpu)))) for w in w0])
    okv.init(keys, [w.copy() for w in w0])
    optname = [None, "sgd", "adam"][seed % 3]
    kw = {"sgd": dict(learning_rate=0.05, momentum=0.9, wd=1e-3), "adam": dict(learning_rate=0.01, wd=1e-3)}.get(optname)
    switch_at = int(rng.integers(0, 4))

    def ctx_of(d):
        return mx.cpu() if d < 0 else mx.gpu(d)

    def want_of(k):
        want = np.empty(shapes[k], np.float32)
        okv.pull(k, want)
        return want

    for step in range(10):
        if optname and step == switch_at:
            kv.set_optimizer(mx.optimizer.create(optname, **kw))
            okv.set_optimizer(O.OracleOptimizer(optname, **kw))
        ks = sorted(int(x) for x in rng.choice(keys, size=int(rng.integers(1, len(keys) + 1)), replace=False))
        devs = [int(x) for x in rng.choice(ngpu, size=int(rng.integers(1, ngpu + 1)), replace=False)]
        if rng.random() < 0.15:
            devs = [-1] + devs[: max(0, len(devs) - 1)]          # one value from the host: no tree for this call
        on_tree = len(devs) >= 3 and min(devs) >= 0
        okv.tree = trees[len(devs)] if on_tree else None
        grads = [[rng.uniform(-1, 1, shapes[k]).astype(np.float32) for _ in devs] for k in ks]
        vals = [[mx.nd.array(g, ctx_of(d)) for g, d in zip(gs, devs)] for gs in grads]
        before = mx.kv.launch_count("tree")
        if rng.random() < 0.5:
            kv.push(ks, vals)
            okv.push(ks, grads)
        else:
            odevs = [int(x) for x in rng.choice(ngpu, size=int(rng.integers(1, ngpu + 1)), replace=False)]
            outs = [[mx.nd.empty(shapes[k], mx.gpu(d)) for d in odevs] for k in ks]
            kv.pushpull(ks, vals, out=outs)
            okv.push(ks, grads)
            for k, oo in zip(ks, outs):
                want = want_of(k)
                for o in oo:
                    assert _bits_equal(o.asnumpy(), want), ("pushpull", seed, step, k, devs)
        assert (mx.kv.launch_count("tree") > before) == on_tree, (seed, step, devs)
        for k in ks:
            d = int(rng.integers(-1, ngpu))
            o = mx.nd.empty(shapes[k], ctx_of(d))
            kv.pull(k, out=o)
            assert _bits_equal(o.asnumpy(), want_of(k)), ("pull", seed, step, k, d, devs)


@pytest.mark.parametrize("world", [3, 4, 8])
def test_one_process_per_gpu_under_the_tree(world):
    """the torchrun deployment shape: tests/mp_worker.py's tree scenario alone (plain and arena-resident arrays, keys
    the ranks shard, SGD-momentum and Adam on the store), every rank checking itself against the oracle."""
    if os.environ.get("MXKV_SIM"):
        pytest.skip("the simulated runtime launches its workers itself (test_sim_host_logic.py)")
    if mx.num_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    env["MXKV_MP_TREE_ONLY"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), os.path.join(root, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("MP_WORKER_OK") == world
