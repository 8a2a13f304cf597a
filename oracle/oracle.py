"""CPU oracle for the KVStore gradient path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``incubator-mxnet_b200``) never does.

Two layers:

* thin ctypes bindings over ``oracle/libkvoracle.so`` (``kv_oracle.c``: the C
  restatement of the reference kernels, each citing its reference file:line) and,
  when present, ``oracle/_ref/libkvref.so`` (the reference's own mshadow
  arithmetic, see ``ref_harness.cc``);
* :class:`OracleKVStore`, a numpy model of ``KVStoreLocal``
  (src/kvstore/kvstore_local.h:70-552) + ``CommDevice`` / ``CommCPU``
  (src/kvstore/comm.h) + the Python ``Updater``/``Optimizer`` bookkeeping
  (python/mxnet/optimizer/updater.py:39-93, optimizer.py:320-352, sgd.py:156-234,
  adam.py:149-186).  Pinned against the reference's known-answer tests in
  tests/test_oracle_golden.py.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u16p = ctypes.POINTER(ctypes.c_uint16)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    """(Re)build the oracle .so (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkvoracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.kvo_adam_lr.restype = ctypes.c_double
        _LIB.kvo_adam_lr.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int]
        _LIB.kvo_unique_i64.restype = ctypes.c_int64
        _LIB.kvo_rsp_sum_f32.restype = ctypes.c_int64
        _LIB.kvo_sum_sq_f32.restype = ctypes.c_float
        _LIB.kvo_sum_sq_f64.restype = ctypes.c_double
        _LIB.kvo_count_nonfinite_f32.restype = ctypes.c_int64
        _LIB.kvo_lars_ratio_f32.restype = ctypes.c_float
    return _LIB


def ref_lib():
    """The reference-arithmetic harness, or None when it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libkvref.so")
        if os.path.exists(path):
            _REF = ctypes.CDLL(path)
    return _REF


def _ptr(a, ct):
    return a.ctypes.data_as(ct)


def _ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


_F = ctypes.c_float
_I64 = ctypes.c_int64

_SUM_DEVICE = {
    np.dtype(np.float32): "kvo_sum_device_f32", np.dtype(np.float64): "kvo_sum_device_f64",
    np.dtype(np.int32): "kvo_sum_device_i32", np.dtype(np.int64): "kvo_sum_device_i64",
    np.dtype(np.uint8): "kvo_sum_device_u8", np.dtype(np.int8): "kvo_sum_device_i8",
    np.dtype(np.float16): "kvo_sum_device_f16",
}


def sum_device(srcs, bf16=False):
    """CommDevice dense reduce order (comm.h:525-548 + ndarray_function-inl.h:457-486).
    ``bf16=True``: inputs are uint16 bf16 bit patterns (fp32 accumulate, one RNE)."""
    srcs = [np.ascontiguousarray(s) for s in srcs]
    out = np.empty_like(srcs[0])
    if bf16:
        assert srcs[0].dtype == np.uint16
        fn = lib().kvo_sum_device_bf16
    else:
        fn = getattr(lib(), _SUM_DEVICE[srcs[0].dtype])
    fn(ctypes.c_int(len(srcs)), _ptr_array(srcs), _I64(srcs[0].size), ctypes.c_void_p(out.ctypes.data))
    return out


def sum_device_lp_f32out(srcs, kind):
    """fp32 sum of fp16 (kind=1) / bf16 (kind=2) inputs given as uint16 bit patterns."""
    srcs = [np.ascontiguousarray(s).view(np.uint16) for s in srcs]
    out = np.empty(srcs[0].shape, np.float32)
    fn = lib().kvo_sum_device_f16_f32out if kind == 1 else lib().kvo_sum_device_bf16_f32out
    fn(ctypes.c_int(len(srcs)), _ptr_array(srcs), _I64(srcs[0].size), _ptr(out, c_f32p))
    return out


def sum_cpu(srcs, nthreads=4, bigarray_bound=1000 * 1000):
    """CommCPU reduce (comm.h:359-411).  Returns a new array (srcs[0] is copied first,
    like CopyFromTo(src[0], &buf_merged))."""
    bufs = [np.array(s, dtype=np.float32, copy=True, order="C") for s in srcs]
    lib().kvo_sum_cpu_f32(ctypes.c_int(len(bufs)), _ptr_array(bufs), _I64(bufs[0].size),
                          ctypes.c_int(nthreads), _I64(bigarray_bound))
    return bufs[0]


def sum_cpu_inplace(bufs, nthreads=4, bigarray_bound=1000 * 1000):
    """In-place CommCPU reduce into bufs[0] (no staging copy) -- the timed kernel of
    the CPU baseline."""
    lib().kvo_sum_cpu_f32(ctypes.c_int(len(bufs)), _ptr_array(bufs), _I64(bufs[0].size),
                          ctypes.c_int(nthreads), _I64(bigarray_bound))


# ---------------------------------------------------------------------------
# MXNET_KVSTORE_USETREE=1: CommDeviceTree (src/kvstore/comm_tree.h)
# ---------------------------------------------------------------------------
def tree_reduce(srcs, topo_row, scan_row, depth, add=None):
    """CommDeviceTree::ReduceInner (comm_tree.h:91-177), level by level on per-GPU merge buffers.

    ``srcs[g]`` is GPU g's value; ``topo_row`` / ``scan_row`` one tree in the reference's array form.  Every GPU of the
    leaf level copies its value into its merge buffer (:108-121); from the deepest level up, nodes are visited in
    pairs (dest, from): a `from` different from `dest` is copied into dest's receive buffer (:137-145), then every
    parent whose two children differ sets its merge buffer to (own buffer + receive buffer) (:151-166).  Returns the
    root's buffer."""
    add = add or (lambda a, b: a + b)
    n = len(srcs)
    topo = [int(v) for v in topo_row]
    scan = [int(v) for v in scan_row]
    merged = [None] * n
    for j in range(scan[depth], scan[depth + 1]):
        g = topo[j]
        merged[g] = np.array(srcs[g], copy=True)
    for level in range(depth, 0, -1):
        reduce_lists = [[] for _ in range(n)]
        copy_buf = [None] * n
        is_dest, dest_id = 0, 0
        for j in range(scan[level], scan[level + 1]):
            topo_id = topo[j]
            if is_dest == 0:
                dest_id = topo_id
                if not reduce_lists[dest_id]:
                    reduce_lists[dest_id].append(("merged", dest_id))
            elif dest_id != topo_id:
                assert copy_buf[dest_id] is None, "two sends into one receive buffer"
                copy_buf[dest_id] = merged[topo_id].copy()
                reduce_lists[dest_id].append(("copy", dest_id))
            is_dest = 0 if is_dest == 1 else 1          # kBranch = 2
        source = scan[level]
        for i in range(scan[level - 1], scan[level]):
            gpu_id = topo[i]
            dest, frm = topo[source], topo[source + 1]
            source += 2
            if len(reduce_lists[gpu_id]) > 1 and dest != frm:
                acc = None
                for kind, g in reduce_lists[gpu_id]:
                    v = merged[g] if kind == "merged" else copy_buf[g]
                    acc = v if acc is None else add(acc, v)
                merged[gpu_id] = acc
    return merged[topo[0]]


def sum_tree(srcs, topo, scan, depth, bound=10000000, add=None):
    """CommDeviceTree::Reduce + Broadcast (comm_tree.h:179-250, :283-325) as seen by a pull: keys above ``bound``
    elements whose first dimension is at least 2n are cut into n row slices (slice_size = rows // n, the last slice
    takes the remainder), slice i reduced up the tree rooted at GPU i; everything else goes up tree 0 whole."""
    n = len(srcs)
    shape = srcs[0].shape
    rows = shape[0] if len(shape) else 1
    if srcs[0].size > bound and rows >= 2 * n:
        out = np.empty_like(srcs[0])
        step = rows // n
        for i in range(n):
            lo, hi = i * step, (rows if i == n - 1 else (i + 1) * step)
            out[lo:hi] = tree_reduce([s[lo:hi] for s in srcs], topo[i], scan[i], depth, add)
        return out
    return tree_reduce(srcs, topo[0], scan[0], depth, add)


def ref_topology_lib():
    """oracle/_ref/libkvref_topo.so: the reference's own gpu_topology.h compiled in place (oracle/ref_topology.cc)."""
    path = os.path.join(_HERE, "_ref", "libkvref_topo.so")
    return ctypes.CDLL(path) if os.path.exists(path) else None


def ref_compute_trees(W, alpha=0.7, backtrack=False):
    """ComputeTrees of the reference (gpu_topology.h:1111-1157) on link matrix W -> (topo [n, 2^(d+1)-1],
    scan [n, d+2]); None where the reference aborts."""
    lib_ = ref_topology_lib()
    W = np.ascontiguousarray(W, np.float32)
    n = W.shape[0]
    topo = np.zeros(n * 64, np.uint64)
    scan = np.zeros(n * 16, np.uint64)
    tl, sl = ctypes.c_int(), ctypes.c_int()
    rc = lib_.kvref_topo_compute_trees(ctypes.c_void_p(W.ctypes.data), ctypes.c_int(n), ctypes.c_float(alpha),
                                       ctypes.c_int(1 if backtrack else 0), ctypes.c_void_p(topo.ctypes.data),
                                       ctypes.c_void_p(scan.ctypes.data), ctypes.byref(tl), ctypes.byref(sl))
    if rc == -2:
        return None
    assert rc == 0
    return (topo[:n * tl.value].reshape(n, tl.value).astype(np.int64),
            scan[:n * sl.value].reshape(n, sl.value).astype(np.int64))


def ref_sum_device(srcs):
    r = ref_lib()
    srcs = [np.ascontiguousarray(s) for s in srcs]
    out = np.empty_like(srcs[0])
    name = {np.dtype(np.float32): "kvref_sum_device_f32", np.dtype(np.float64): "kvref_sum_device_f64",
            np.dtype(np.int32): "kvref_sum_device_i32", np.dtype(np.float16): "kvref_sum_device_f16"}[srcs[0].dtype]
    getattr(r, name)(ctypes.c_int(len(srcs)), _ptr_array(srcs), _I64(srcs[0].size),
                     ctypes.c_void_p(out.ctypes.data))
    return out


def ref_sum_cpu(srcs, nthreads=4, bigarray_bound=1000 * 1000):
    bufs = [np.array(s, dtype=np.float32, copy=True, order="C") for s in srcs]
    ref_lib().kvref_sum_cpu_f32(ctypes.c_int(len(bufs)), _ptr_array(bufs), _I64(bufs[0].size),
                                ctypes.c_int(nthreads), _I64(bigarray_bound))
    return bufs[0]


# ---------------------------------------------------------------------------
# bf16 helpers (bit patterns in uint16)
# ---------------------------------------------------------------------------
def f32_to_bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


def bf16_to_f32(h):
    return (np.ascontiguousarray(h, np.uint16).astype(np.uint32) << 16).view(np.float32)


# ---------------------------------------------------------------------------
# optimizer kernels (in place on numpy arrays)
# ---------------------------------------------------------------------------
def _clip(c):
    return _F(-1.0 if c is None else c)


def sgd_update(w, g, lr, wd=0.0, rescale=1.0, clip=None):
    lib().kvo_sgd_update_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(w, c_f32p), _ptr(g, c_f32p),
                             _F(lr), _F(wd), _F(rescale), _clip(clip))


def sgd_mom_update(w, g, mom, lr, wd=0.0, momentum=0.0, rescale=1.0, clip=None):
    lib().kvo_sgd_mom_update_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(mom, c_f32p), _ptr(w, c_f32p),
                                 _ptr(g, c_f32p), _F(lr), _F(wd), _F(momentum), _F(rescale), _clip(clip))


def mp_sgd_update(w_lp, lp_kind, w32, g32, lr, wd=0.0, rescale=1.0, clip=None):
    lib().kvo_mp_sgd_update(_I64(w32.size), ctypes.c_void_p(w_lp.ctypes.data if w_lp is not None else 0),
                            ctypes.c_int(lp_kind), _ptr(w32, c_f32p), _ptr(g32, c_f32p),
                            _F(lr), _F(wd), _F(rescale), _clip(clip))


def mp_sgd_mom_update(w_lp, lp_kind, w32, mom, g32, lr, wd=0.0, momentum=0.0, rescale=1.0, clip=None):
    lib().kvo_mp_sgd_mom_update(_I64(w32.size), ctypes.c_void_p(w_lp.ctypes.data if w_lp is not None else 0),
                                ctypes.c_int(lp_kind), _ptr(w32, c_f32p), _ptr(mom, c_f32p),
                                _ptr(g32, c_f32p), _F(lr), _F(wd), _F(momentum), _F(rescale), _clip(clip))


def adam_update(w, g, mean, var, lr, wd=0.0, beta1=0.9, beta2=0.999, eps=1e-8, rescale=1.0, clip=None):
    lib().kvo_adam_update_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(mean, c_f32p), _ptr(var, c_f32p),
                              _ptr(w, c_f32p), _ptr(g, c_f32p), _F(lr), _F(wd), _F(beta1), _F(beta2),
                              _F(eps), _F(rescale), _clip(clip))


def mp_adam_update(w_lp, lp_kind, w32, mean, var, g32, lr, wd=0.0, beta1=0.9, beta2=0.999, eps=1e-8,
                   rescale=1.0, clip=None):
    lib().kvo_mp_adam_update(_I64(w32.size), ctypes.c_void_p(w_lp.ctypes.data if w_lp is not None else 0),
                             ctypes.c_int(lp_kind), _ptr(w32, c_f32p), _ptr(mean, c_f32p),
                             _ptr(var, c_f32p), _ptr(g32, c_f32p), _F(lr), _F(wd), _F(beta1), _F(beta2),
                             _F(eps), _F(rescale), _clip(clip))


def mp_adamw_update(w_lp, lp_kind, w32, mean, var, g32, lr, eta=1.0, wd=0.0, beta1=0.9, beta2=0.999,
                    eps=1e-8, rescale=1.0, clip=None):
    lib().kvo_mp_adamw_update(_I64(w32.size), ctypes.c_void_p(w_lp.ctypes.data if w_lp is not None else 0),
                              ctypes.c_int(lp_kind), _ptr(w32, c_f32p), _ptr(mean, c_f32p),
                              _ptr(var, c_f32p), _ptr(g32, c_f32p), _F(lr), _F(eta), _F(wd), _F(beta1),
                              _F(beta2), _F(eps), _F(rescale), _clip(clip))


# ---- layer-wise adaptive optimizers (multi_sum_sq / multi_lamb / multi_lans / LARS) -------------
def sum_sq(x, scale=1.0, mode="seq"):
    """Sum of squares as a float32.  mode 'seq': the reference CPU operator's sequential float sum
    (multi_sum_sq.cc:42-62); 'f64': accumulated in double and rounded once -- the value every float
    summation order (CPU sequential, GPU block tree) approximates."""
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    if mode == "seq":
        return np.float32(lib().kvo_sum_sq_f32(_I64(x.size), _ptr(x, c_f32p), _F(scale)))
    return np.float32(lib().kvo_sum_sq_f64(_I64(x.size), _ptr(x, c_f32p), _F(scale)))


def count_nonfinite(x):
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    return int(lib().kvo_count_nonfinite_f32(_I64(x.size), _ptr(x, c_f32p)))


def lamb_update(w, g, mean, var, lr, wd, t, beta1=0.9, beta2=0.999, eps=1e-6, rescale=1.0, clip=None,
                bias_correction=True, lower_bound=None, upper_bound=None, norm_mode="seq"):
    """multi_lamb_update on one tensor (multi_lamb-inl.h:248-330): r1 from the weight, step 1, r2 from
    the update direction, step 2.  ``w`` (float32, the master in the mp variant) is updated in place."""
    tmp = np.empty_like(w)
    ssw = sum_sq(w, 1.0, norm_mode)
    lib().kvo_lamb_step1_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(g, c_f32p), _ptr(mean, c_f32p), _ptr(var, c_f32p),
                             _ptr(tmp, c_f32p), _F(beta1), _F(beta2), _F(eps), _F(wd), _F(rescale), _clip(clip),
                             ctypes.c_int(1 if bias_correction else 0), ctypes.c_int(t))
    ssg = sum_sq(tmp, 1.0, norm_mode)
    lib().kvo_lamb_step2_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(tmp, c_f32p), _F(lr), _F(ssw), _F(ssg),
                             _F(lower_bound if lower_bound else -1.0), _F(upper_bound if upper_bound else -1.0))


def lans_update(w, g, mean, var, lr, wd, t, beta1=0.9, beta2=0.999, eps=1e-6, rescale=1.0, clip=None,
                lower_bound=None, upper_bound=None, norm_mode="seq", w_norm_src=None):
    """multi_lans_update on one tensor (multi_lans-inl.h:262-380).  ``w_norm_src``: the array whose norm
    is r1 -- the reference takes the low-precision weight even in the mp variant (:296-300)."""
    tm, tg = np.empty_like(w), np.empty_like(w)
    ssw = sum_sq(w if w_norm_src is None else w_norm_src, 1.0, norm_mode)
    gsq = sum_sq(g, rescale, norm_mode)
    lib().kvo_lans_step1_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(g, c_f32p), _ptr(mean, c_f32p), _ptr(var, c_f32p),
                             _ptr(tm, c_f32p), _ptr(tg, c_f32p), _F(beta1), _F(beta2), _F(eps), _F(wd), _F(rescale),
                             _clip(clip), ctypes.c_int(t), _F(gsq))
    ssm, ssg = sum_sq(tm, 1.0, norm_mode), sum_sq(tg, 1.0, norm_mode)
    lib().kvo_lans_step2_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(tm, c_f32p), _ptr(tg, c_f32p), _F(lr), _F(beta1),
                             _F(ssw), _F(ssm), _F(ssg), _F(lower_bound if lower_bound else -1.0),
                             _F(upper_bound if upper_bound else -1.0))


def lars_lr(lr, w_for_norm, g, wd, eta=0.001, eps=1e-8, rescale=1.0, norm_mode="seq"):
    """LARS._get_lars (lars.py:117-133) folded into the learning rate like lars.py:258-260: the
    Python double ``lr`` times the float32 ratio."""
    ssw = sum_sq(w_for_norm, 1.0, norm_mode)
    # `v *= self.rescale_grad` is always applied (lars.py:108-110)
    gs = np.ascontiguousarray(g, dtype=np.float32) * np.float32(rescale)
    ssg = sum_sq(gs, 1.0, norm_mode)
    lars = lib().kvo_lars_ratio_f32(_F(ssw), _F(ssg), _F(eta), _F(wd), _F(eps))
    return float(lr) * float(lars)


def test_update(w, g, lr, wd=0.0, rescale=1.0):
    lib().kvo_test_update_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(g, c_f32p), _F(lr), _F(wd), _F(rescale))


def adam_lr(lr, beta1, beta2, t):
    return lib().kvo_adam_lr(lr, beta1, beta2, t)


# ---------------------------------------------------------------------------
# row-sparse
# ---------------------------------------------------------------------------
class RowSparse(object):
    """(indices int64 sorted unique [nnz], data [nnz, L...], shape) -- MXNet's
    row_sparse NDArray (include/mxnet/ndarray.h:61-66)."""

    def __init__(self, indices, data, shape):
        self.indices = np.ascontiguousarray(indices, np.int64)
        self.data = np.ascontiguousarray(data, np.float32)
        self.shape = tuple(shape)

    @staticmethod
    def from_dense(a):
        a = np.asarray(a, np.float32)
        nz = np.where(np.any(a.reshape(a.shape[0], -1) != 0, axis=1))[0]
        return RowSparse(nz, a[nz], a.shape)

    def todense(self):
        out = np.zeros(self.shape, np.float32)
        if len(self.indices):
            out[self.indices] = self.data.reshape((len(self.indices),) + self.shape[1:])
        return out

    def copy(self):
        return RowSparse(self.indices.copy(), self.data.copy(), self.shape)


def unique(ids):
    a = np.array(ids, dtype=np.int64, copy=True).reshape(-1)
    n = lib().kvo_unique_i64(_ptr(a, c_i64p), _I64(a.size))
    return a[:n].copy()


def rsp_sum(rsps):
    L = int(np.prod(rsps[0].shape[1:]))
    n = len(rsps)
    tot = sum(len(r.indices) for r in rsps)
    out_idx = np.empty(max(tot, 1), np.int64)
    out_val = np.empty((max(tot, 1), L), np.float32)
    idxs = [r.indices for r in rsps]
    vals = [r.data.reshape(-1, L) if len(r.indices) else np.zeros((0, L), np.float32) for r in rsps]
    vals = [np.ascontiguousarray(v) for v in vals]
    nnz = np.array([len(r.indices) for r in rsps], np.int64)
    nu = lib().kvo_rsp_sum_f32(ctypes.c_int(n), _ptr_array(idxs), _ptr_array(vals), _ptr(nnz, c_i64p),
                               _I64(L), _ptr(out_idx, c_i64p), _ptr(out_val, c_f32p))
    return RowSparse(out_idx[:nu].copy(), out_val[:nu].copy(), rsps[0].shape)


def sparse_retain(src, ids):
    L = int(np.prod(src.shape[1:]))
    ids = np.ascontiguousarray(ids, np.int64).reshape(-1)
    out_idx = np.empty(max(len(ids), 1), np.int64)
    out_val = np.empty((max(len(ids), 1), L), np.float32)
    sval = np.ascontiguousarray(src.data.reshape(-1, L))
    lib().kvo_sparse_retain_f32(_ptr(src.indices, c_i64p), _ptr(sval, c_f32p), _I64(len(src.indices)),
                                _I64(L), _ptr(ids, c_i64p), _I64(len(ids)), _ptr(out_idx, c_i64p),
                                _ptr(out_val, c_f32p))
    return RowSparse(out_idx[:len(ids)].copy(), out_val[:len(ids)].copy(), src.shape)


def sgd_rsp_lazy(w, grad, lr, wd=0.0, rescale=1.0, clip=None, mom=None, momentum=0.0):
    L = int(np.prod(w.shape[1:]))
    gv = np.ascontiguousarray(grad.data.reshape(-1, L))
    if mom is None:
        lib().kvo_sgd_rsp_lazy_f32(_ptr(w, c_f32p), _I64(L), _ptr(grad.indices, c_i64p), _ptr(gv, c_f32p),
                                   _I64(len(grad.indices)), _F(lr), _F(wd), _F(rescale), _clip(clip))
    else:
        lib().kvo_sgd_mom_rsp_lazy_f32(_ptr(w, c_f32p), _ptr(mom, c_f32p), _I64(L),
                                       _ptr(grad.indices, c_i64p), _ptr(gv, c_f32p),
                                       _I64(len(grad.indices)), _F(lr), _F(wd), _F(momentum), _F(rescale),
                                       _clip(clip))


def sgd_std_rsp(w, grad, lr, wd=0.0, rescale=1.0, clip=None):
    L = int(np.prod(w.shape[1:]))
    gv = np.ascontiguousarray(grad.data.reshape(-1, L))
    lib().kvo_sgd_std_rsp_f32(_ptr(w, c_f32p), _I64(w.shape[0]), _I64(L), _ptr(grad.indices, c_i64p),
                              _ptr(gv, c_f32p), _I64(len(grad.indices)), _F(lr), _F(wd), _F(rescale), _clip(clip))


def adam_std_update(w, gdense, mean, var, lr, wd=0.0, beta1=0.9, beta2=0.999, eps=1e-8, rescale=1.0, clip=None):
    lib().kvo_adam_std_update_f32(_I64(w.size), _ptr(w, c_f32p), _ptr(mean, c_f32p), _ptr(var, c_f32p),
                                  _ptr(gdense, c_f32p), _F(lr), _F(wd), _F(beta1), _F(beta2), _F(eps),
                                  _F(rescale), _clip(clip))


def adam_rsp_lazy(w, grad, mean, var, lr, wd=0.0, beta1=0.9, beta2=0.999, eps=1e-8, rescale=1.0, clip=None):
    L = int(np.prod(w.shape[1:]))
    gv = np.ascontiguousarray(grad.data.reshape(-1, L))
    lib().kvo_adam_rsp_lazy_f32(_ptr(w, c_f32p), _ptr(mean, c_f32p), _ptr(var, c_f32p), _I64(L),
                                _ptr(grad.indices, c_i64p), _ptr(gv, c_f32p), _I64(len(grad.indices)),
                                _F(lr), _F(wd), _F(beta1), _F(beta2), _F(eps), _F(rescale), _clip(clip))


# ---------------------------------------------------------------------------
# gradient compression
# ---------------------------------------------------------------------------
def quantize_2bit(grad, residual, thr):
    E = grad.size
    out = np.zeros(((E + 15) // 16) * 4, np.uint8)
    lib().kvo_quantize_2bit(_I64(E), _ptr(grad, c_f32p), _ptr(residual, c_f32p), _ptr(out, c_u8p), _F(thr))
    return out


def dequantize_2bit(comp, E, thr):
    out = np.empty(E, np.float32)
    lib().kvo_dequantize_2bit(_I64(E), _ptr(comp, c_u8p), _ptr(out, c_f32p), _F(thr))
    return out


def quantize_1bit(grad, residual, thr):
    E = grad.size
    out = np.zeros(((E + 31) // 32) * 4, np.uint8)
    lib().kvo_quantize_1bit(_I64(E), _ptr(grad, c_f32p), _ptr(residual, c_f32p), _ptr(out, c_u8p), _F(thr))
    return out


def dequantize_1bit(comp, E, thr):
    out = np.empty(E, np.float32)
    lib().kvo_dequantize_1bit(_I64(E), _ptr(comp, c_u8p), _ptr(out, c_f32p), _F(thr))
    return out


# ---------------------------------------------------------------------------
# Optimizer bookkeeping (python/mxnet/optimizer/optimizer.py) + Updater
# ---------------------------------------------------------------------------
class OracleOptimizer(object):
    """Hyper-parameter bookkeeping of mx.optimizer.Optimizer: per-index update
    counts (optimizer.py:_update_count), lr/wd multipliers (_get_lr/_get_wd),
    rescale_grad, clip_gradient, multi_precision."""

    def __init__(self, name, learning_rate=None, wd=0.0, rescale_grad=1.0, clip_gradient=None,
                 momentum=0.0, beta1=0.9, beta2=0.999, epsilon=None, eta=None, multi_precision=False,
                 lr_mult=None, wd_mult=None, lazy_update=False, lower_bound=None, upper_bound=None,
                 bias_correction=True, norm_mode="seq", no_trust=(), correct_bias=True, begin_num_update=0):
        self.name = name.lower()
        self.begin_num_update = begin_num_update     # optimizer.py:111-112, 445-462
        self.lower_bound, self.upper_bound, self.bias_correction = lower_bound, upper_bound, bias_correction
        self.norm_mode = norm_mode
        self.correct_bias = correct_bias    # AdamW (adamW.py:80-88)
        self.no_trust = set(no_trust)       # LARS: indices named *gamma / *beta / *bias
        if learning_rate is None:
            # sgd.py:95, adam.py:85, adamW.py:80, lamb.py:66, lans.py:62, lars.py:77; optimizer.py:100-101 otherwise
            learning_rate = {"sgd": 0.1, "adam": 0.001, "adamw": 0.001, "lamb": 0.001, "lans": 0.001,
                             "lars": 0.1}.get(self.name, 0.01)
        self.lr = learning_rate
        self.wd = wd
        self.rescale_grad = rescale_grad
        self.clip_gradient = clip_gradient
        self.momentum = momentum
        if epsilon is None:      # adam.py:85, lars.py:78: 1e-8; adamW.py:87, lamb.py:67, lans.py:62: 1e-6
            epsilon = 1e-6 if self.name in ("adamw", "lamb", "lans") else 1e-8
        if eta is None:          # AdamW's schedule multiplier (1.0) / LARS' trust coefficient (0.001)
            eta = 0.001 if self.name == "lars" else 1.0
        self.beta1, self.beta2, self.epsilon, self.eta = beta1, beta2, epsilon, eta
        self.multi_precision = multi_precision
        self.lr_mult = dict(lr_mult or {})
        self.wd_mult = dict(wd_mult or {})
        self.lazy_update = lazy_update
        self.count = {}
        self.states = {}

    def _lr(self, index):
        return self.lr * self.lr_mult.get(index, 1.0)

    def _wd(self, index):
        return self.wd * self.wd_mult.get(index, 1.0)

    def update(self, index, weight, grad):
        """Updater.__call__ (updater.py:39-93): create state on first sight, then one
        fused step.  ``weight`` is a float32 ndarray updated in place; ``grad`` is a
        float32 ndarray or RowSparse."""
        self.count[index] = t = self.count.get(index, self.begin_num_update) + 1
        lr, wd = self._lr(index), self._wd(index)
        n = self.name
        sparse = isinstance(grad, RowSparse)
        if n == "test":
            test_update(weight, grad if not sparse else grad.todense(), lr, wd, self.rescale_grad)
        elif n == "sgd":
            if self.momentum != 0.0 and index not in self.states:
                self.states[index] = np.zeros_like(weight)
            mom = self.states.get(index)
            if sparse and not self.lazy_update:
                # standard update (sgd.py default lazy_update=False): optimizer_op-inl.h:471-515 /
                # SGDMomStdDnsRspDnsKernel optimizer_op.cu:32-57 (= the dense kernel on the densified grad)
                if mom is None:
                    sgd_std_rsp(weight, grad, lr, wd, self.rescale_grad, self.clip_gradient)
                else:
                    sgd_mom_update(weight, np.ascontiguousarray(grad.todense()), mom, lr, wd, self.momentum,
                                   self.rescale_grad, self.clip_gradient)
            elif sparse:
                sgd_rsp_lazy(weight, grad, lr, wd, self.rescale_grad, self.clip_gradient, mom, self.momentum)
            elif mom is None:
                sgd_update(weight, grad, lr, wd, self.rescale_grad, self.clip_gradient)
            else:
                sgd_mom_update(weight, grad, mom, lr, wd, self.momentum, self.rescale_grad, self.clip_gradient)
        elif n == "adam":
            if index not in self.states:
                self.states[index] = (np.zeros_like(weight), np.zeros_like(weight))
            mean, var = self.states[index]
            lr = adam_lr(lr, self.beta1, self.beta2, t)
            if sparse and not self.lazy_update:
                adam_std_update(weight, np.ascontiguousarray(grad.todense()), mean, var, lr, wd, self.beta1,
                                self.beta2, self.epsilon, self.rescale_grad, self.clip_gradient)
            elif sparse:
                adam_rsp_lazy(weight, grad, mean, var, lr, wd, self.beta1, self.beta2, self.epsilon,
                              self.rescale_grad, self.clip_gradient)
            else:
                adam_update(weight, grad, mean, var, lr, wd, self.beta1, self.beta2, self.epsilon,
                            self.rescale_grad, self.clip_gradient)
        elif n == "adamw":
            if index not in self.states:
                self.states[index] = (np.zeros_like(weight), np.zeros_like(weight))
            mean, var = self.states[index]
            # AdamW.fused_step hands the operator lr = 1 and eta = the (bias-corrected) learning rate
            # (`lrs=np.ones(...)`, `etas=lrs`, adamW.py:176-200); self.eta is the engine's extra multiplier
            lr_c = adam_lr(lr, self.beta1, self.beta2, t) if self.correct_bias else lr
            rg = self.rescale_grad
            if not np.isfinite(rg) or rg == 0:      # the operator skips the update (adamw-inl.h:455)
                return
            mp_adamw_update(None, 0, weight, mean, var, grad, 1.0, float(np.float32(lr_c * self.eta)), wd,
                            self.beta1, self.beta2, self.epsilon, rg, self.clip_gradient)
        elif n in ("lamb", "lans"):
            assert not sparse
            if index not in self.states:
                self.states[index] = (np.zeros_like(weight), np.zeros_like(weight))
            mean, var = self.states[index]
            kw = dict(beta1=self.beta1, beta2=self.beta2, eps=self.epsilon, rescale=self.rescale_grad,
                      clip=self.clip_gradient, lower_bound=self.lower_bound, upper_bound=self.upper_bound,
                      norm_mode=self.norm_mode)
            if n == "lamb":
                lamb_update(weight, grad, mean, var, lr, wd, t, bias_correction=self.bias_correction, **kw)
            else:
                lans_update(weight, grad, mean, var, lr, wd, t, **kw)
        elif n == "lars":
            assert not sparse
            if index not in self.no_trust:
                lr = lars_lr(lr, weight, grad, wd, self.eta, self.epsilon, self.rescale_grad, self.norm_mode)
            if self.momentum != 0.0:
                if index not in self.states:
                    self.states[index] = np.zeros_like(weight)
                sgd_mom_update(weight, grad, self.states[index], lr, wd, self.momentum, self.rescale_grad,
                               self.clip_gradient)
            else:
                sgd_update(weight, grad, lr, wd, self.rescale_grad, self.clip_gradient)
        else:
            raise ValueError("unknown optimizer " + n)


class OracleKVStore(object):
    """numpy model of KVStoreLocal (kvstore_local.h).  ``type`` decides the reduce
    association: names containing 'device' use CommDevice order, others CommCPU
    (kvstore.cc:42-85)."""

    def __init__(self, type="local", tree=None):
        self.type = type
        self.device = "device" in type.lower()
        # MXNET_KVSTORE_USETREE=1 (kvstore_local.h:74-82): dict(topo=, scan=, depth=, bound=) of CommDeviceTree
        self.tree = tree if self.device else None
        self.local = {}
        self.key_type = None
        self.updater = None
        self.optimizer = None
        self.nthreads = 4

    # -- helpers -----------------------------------------------------------
    def _set_key_type(self, k):
        kt = str if isinstance(k, str) else int
        if self.key_type is None:
            self.key_type = kt
        if self.key_type is not kt:
            raise ValueError("Mixed key types are not allowed")  # kvstore_local.h:344-347

    @staticmethod
    def _flatten(keys, vals):
        """_ctype_key_value (python/mxnet/kvstore/base.py:32-65): key-major flattening."""
        if isinstance(keys, (list, tuple)):
            assert len(keys) == len(vals)
            ks, vs = [], []
            for k, v in zip(keys, vals):
                k2, v2 = OracleKVStore._flatten(k, v)
                ks += k2
                vs += v2
            return ks, vs
        if isinstance(vals, (np.ndarray, RowSparse)):
            return [keys], [vals]
        return [keys] * len(vals), list(vals)

    @staticmethod
    def _group(keys, vals):
        """GroupKVPairs (kvstore_local.h:440-469): sort by key, group equal keys."""
        order = sorted(range(len(keys)), key=lambda i: keys[i])
        uniq, grouped = [], []
        for i in order:
            if not uniq or keys[i] != uniq[-1]:
                uniq.append(keys[i])
                grouped.append([vals[i]])
            else:
                grouped[-1].append(vals[i])
        return uniq, grouped

    def _reduce(self, vals):
        if len(vals) == 1:            # comm.h:128-131 / :514-516
            return vals[0]
        if isinstance(vals[0], RowSparse):
            return rsp_sum(vals)
        if self.device and self.tree is not None and len(vals) == self.tree["topo"].shape[0]:
            t = self.tree
            return sum_tree(vals, t["topo"], t["scan"], t["depth"], t.get("bound", 10000000))
        if self.device:
            return sum_device(vals)
        return sum_cpu(vals, self.nthreads).reshape(vals[0].shape)

    # -- API ---------------------------------------------------------------
    def init(self, key, value):
        ks, vs = self._flatten(key, value)
        for k, v in zip(ks, vs):
            self._set_key_type(k)
            if k in self.local:
                raise ValueError("duplicate init of key %s" % str(k))   # kvstore_local.h:230-233
            self.local[k] = v.copy()

    def set_updater(self, fn):
        self.updater = fn

    def set_optimizer(self, opt):
        self.optimizer = opt
        self.updater = lambda k, recv, local: opt.update(k, local, recv)

    def push(self, key, value, priority=0):
        ks, vs = self._flatten(key, value)
        for k in ks:
            self._set_key_type(k)
        uniq, grouped = self._group(ks, vs)
        for k, g in zip(uniq, grouped):
            merged = self._reduce(g)
            if k not in self.local:
                raise KeyError("key %s has not been inited" % str(k))
            if self.updater is not None:
                local = self.local[k]
                if isinstance(local, RowSparse):
                    dense = local.todense()
                    self.updater(k, merged, dense)
                    self.local[k] = RowSparse.from_dense(dense)
                else:
                    self.updater(k, merged, local)
            else:
                self.local[k] = merged.copy()     # kvstore_local.h:279-284

    def pull(self, key, out, priority=0, ignore_sparse=True):
        ks, vs = self._flatten(key, out)
        for k in ks:
            self._set_key_type(k)
        uniq, grouped = self._group(ks, vs)
        for k, g in zip(uniq, grouped):
            if k not in self.local:
                raise KeyError("key %s has not been inited" % str(k))
            src = self.local[k]
            for o in g:
                if isinstance(o, RowSparse):
                    if ignore_sparse:
                        continue                  # kvstore_local.h:396-407
                    s = src if isinstance(src, RowSparse) else RowSparse.from_dense(src)
                    o.indices, o.data = s.indices.copy(), s.data.copy()
                else:
                    o[...] = src.todense() if isinstance(src, RowSparse) else src

    def pushpull(self, key, value, out=None, priority=0):
        self.push(key, value, priority)
        self.pull(key, out if out is not None else value, priority)

    def broadcast(self, key, value, out, priority=0):
        self.init(key, value)
        self.pull(key, out, priority)

    def row_sparse_pull(self, key, out, row_ids, priority=0):
        ks, vs = self._flatten(key, out)
        _, rs = self._flatten(key, row_ids)
        for k, o, r in zip(ks, vs, rs):
            self._set_key_type(k)
            if not isinstance(o, RowSparse):
                raise ValueError("Expected row_sparse storage type for row_sparse_pull values")
            src = self.local[k]
            if not isinstance(src, RowSparse):
                raise ValueError("PullRowSparse expects row_sparse src NDArray")
            ret = sparse_retain(src, unique(r))
            o.indices, o.data = ret.indices, ret.data.reshape((len(ret.indices),) + src.shape[1:])
