"""Minimal NDArray front-end over the C ABI (the subset of python/mxnet/ndarray/ndarray.py
and sparse.py that the KVStore path and its tests use).  Data lives in memory owned by the
native library (or borrowed from torch); numpy is only a host-side transport here."""
import ctypes
import os

import numpy as np

from .base import _LIB, check_call, MXNetError
from .context import Context, cpu, gpu

# mshadow type flags, 3rdparty/mshadow/mshadow/base.h:352-366
_DTYPE_NP_TO_MX = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.float16): 2,
                   np.dtype(np.uint8): 3, np.dtype(np.int32): 4, np.dtype(np.int8): 5, np.dtype(np.int64): 6}
_DTYPE_MX_TO_NP = {v: k for k, v in _DTYPE_NP_TO_MX.items()}
BFLOAT16 = 12
_STYPE = {"default": 0, "row_sparse": 1}


def _mx_dtype(dtype):
    if isinstance(dtype, str) and dtype in ("bfloat16", "bf16"):
        return BFLOAT16
    if isinstance(dtype, int):
        return dtype
    return _DTYPE_NP_TO_MX[np.dtype(dtype)]


def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _f32_to_bf16(f32):
    u = np.ascontiguousarray(f32, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


class NDArray(object):
    """Handle to a native array.  ``handle`` is the opaque NDArrayHandle."""
    __slots__ = ("handle", "_h", "_keep", "__weakref__")

    def __init__(self, handle, keep=None):
        self.handle = handle if isinstance(handle, ctypes.c_void_p) else ctypes.c_void_p(handle)
        self._h = self.handle.value      # the raw address, read once (marshalling hundreds of arrays per step)
        self._keep = keep     # objects that own borrowed memory (e.g. a torch tensor)

    def __del__(self):
        try:
            if self.handle:
                _LIB.MXNDArrayFree(self.handle)
        except Exception:  # interpreter shutdown
            pass

    # -- metadata ----------------------------------------------------------
    @property
    def shape(self):
        ndim = ctypes.c_int()
        pdata = ctypes.POINTER(ctypes.c_int64)()
        check_call(_LIB.MXNDArrayGetShape64(self.handle, ctypes.byref(ndim), ctypes.byref(pdata)))
        return tuple(int(pdata[i]) for i in range(ndim.value))

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    @property
    def mx_dtype(self):
        t = ctypes.c_int()
        check_call(_LIB.MXNDArrayGetDType(self.handle, ctypes.byref(t)))
        return t.value

    @property
    def dtype(self):
        t = self.mx_dtype
        return "bfloat16" if t == BFLOAT16 else _DTYPE_MX_TO_NP[t].type

    @property
    def context(self):
        dt, di = ctypes.c_int(), ctypes.c_int()
        check_call(_LIB.MXNDArrayGetContext(self.handle, ctypes.byref(dt), ctypes.byref(di)))
        return Context(Context.devtype2str[dt.value], di.value)

    ctx = context

    @property
    def stype(self):
        s = ctypes.c_int()
        check_call(_LIB.MXNDArrayGetStorageType(self.handle, ctypes.byref(s)))
        return {0: "default", 1: "row_sparse", 2: "csr"}.get(s.value, "undefined")

    @property
    def data_ptr(self):
        p = ctypes.c_void_p()
        check_call(_LIB.MXNDArrayGetData(self.handle, ctypes.byref(p)))
        return p.value or 0

    # -- sync ----------------------------------------------------------------
    def wait_to_read(self):
        check_call(_LIB.MXNDArrayWaitToRead(self.handle))

    # -- host transport --------------------------------------------------------
    def asnumpy(self, raw=False):
        if self.stype == "row_sparse":
            return self.todense_numpy()
        t = self.mx_dtype
        npdt = np.dtype(np.uint16) if t == BFLOAT16 else _DTYPE_MX_TO_NP[t]
        out = np.empty(self.shape, npdt)
        check_call(_LIB.MXNDArraySyncCopyToCPU(self.handle, out.ctypes.data_as(ctypes.c_void_p),
                                               ctypes.c_size_t(out.size)))
        if t == BFLOAT16 and not raw:
            return _bf16_to_f32(out)
        return out

    def _sync_copyfrom(self, arr):
        t = self.mx_dtype
        if t == BFLOAT16:
            arr = np.asarray(arr)
            src = arr if arr.dtype == np.uint16 else _f32_to_bf16(arr.astype(np.float32))
        else:
            src = np.ascontiguousarray(arr, _DTYPE_MX_TO_NP[t])
        src = np.ascontiguousarray(src)
        if src.size != self.size:
            src = np.ascontiguousarray(np.broadcast_to(src, self.shape))
        check_call(_LIB.MXNDArraySyncCopyFromCPU(self.handle, src.ctypes.data_as(ctypes.c_void_p),
                                                 ctypes.c_size_t(src.size)))

    def __setitem__(self, key, value):
        if not (isinstance(key, slice) and key == slice(None)):
            raise NotImplementedError("only a[:] = value is supported")
        if isinstance(value, NDArray):
            value.copyto(self)
        elif np.isscalar(value):
            self._sync_copyfrom(np.full(self.shape, value))
        else:
            self._sync_copyfrom(np.asarray(value))

    def copyto(self, other):
        if isinstance(other, Context):
            out = empty(self.shape, other, self.dtype)
            self.copyto(out)
            return out
        if self.stype == "row_sparse":
            if other.stype == "row_sparse":       # values first, then indices (the row count follows the source)
                check_call(_LIB.MXNDArraySyncCopyFromNDArray(other.handle, self.handle, ctypes.c_int(-1)))
                check_call(_LIB.MXNDArraySyncCopyFromNDArray(other.handle, self.handle, ctypes.c_int(0)))
            else:
                other[:] = self.todense_numpy()
            return other
        check_call(_LIB.MXNDArraySyncCopyFromNDArray(other.handle, self.handle, ctypes.c_int(-1)))
        return other

    def copy(self):
        return self.copyto(self.context)

    def as_in_context(self, ctx):
        return self if ctx == self.context else self.copyto(ctx)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        dims = (ctypes.c_int64 * len(shape))(*shape)
        out = ctypes.c_void_p()
        check_call(_LIB.MXNDArrayReshape64(self.handle, len(shape), dims, ctypes.c_bool(False), ctypes.byref(out)))
        return NDArray(out, keep=self._keep)

    def astype(self, dtype):
        out = empty(self.shape, self.context, dtype)
        out[:] = self.asnumpy()
        return out

    def tostype(self, stype):
        if stype == self.stype:
            return self
        if stype == "row_sparse":
            return row_sparse_array(self.asnumpy(), ctx=self.context, dtype=self.dtype)
        if stype == "default":
            return array(self.asnumpy(), ctx=self.context, dtype=self.dtype)
        raise ValueError("unknown stype " + stype)

    # -- torch views: the "embedding framework" side used by updater callbacks -------------
    def as_torch(self):
        """Zero-copy torch view of a dense array (DLPack, c_api.h:976-1002)."""
        import torch
        from torch.utils import dlpack as _dl
        ptr = ctypes.c_void_p()
        check_call(_LIB.MXNDArrayToDLPack(self.handle, ctypes.byref(ptr)))
        ctypes.pythonapi.PyCapsule_New.restype = ctypes.py_object
        ctypes.pythonapi.PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        cap = ctypes.pythonapi.PyCapsule_New(ptr, b"dltensor", None)
        t = _dl.from_dlpack(cap)
        return t

    def _binary_inplace(self, other, op):
        if os.environ.get("MXKV_SIM"):
            # simulated CUDA runtime (tests/sim): torch cannot share a process with the stand-in
            # libcudart, so the convenience arithmetic goes through numpy
            a = self.asnumpy().astype(np.float64 if self.dtype == np.float64 else np.float32)
            if isinstance(other, NDArray) and other.stype == "row_sparse":
                assert op in ("add_", "sub_"), "row_sparse operand supports += / -= only"
                idx = other.indices.asnumpy()
                val = other.data.asnumpy().reshape((idx.size,) + a.shape[1:])
                np.add.at(a, idx, val if op == "add_" else -val)
            else:
                o = other.asnumpy() if isinstance(other, NDArray) else other
                a = {"add_": a + o, "sub_": a - o, "mul_": a * o}[op]
            self._sync_copyfrom(a)
            return self
        t = self.as_torch()
        if isinstance(other, NDArray) and other.stype == "row_sparse":
            # dense (op)= row_sparse: only the stored rows take part (used by updater callbacks)
            assert op in ("add_", "sub_"), "row_sparse operand supports += / -= only"
            idx = other.indices.as_torch().to(t.device)
            val = other.data.as_torch().to(t.device).reshape((idx.numel(),) + tuple(t.shape[1:]))
            t.index_add_(0, idx, val if op == "add_" else -val)
            return self
        o = other.as_torch().to(t.device) if isinstance(other, NDArray) else other
        getattr(t, op)(o)
        return self

    def __iadd__(self, other):
        return self._binary_inplace(other, "add_")

    def __isub__(self, other):
        return self._binary_inplace(other, "sub_")

    def __imul__(self, other):
        return self._binary_inplace(other, "mul_")

    def _binary(self, other, op):
        out = self.copy()
        return out._binary_inplace(other, op)

    def __add__(self, other):
        return self._binary(other, "add_")

    def __sub__(self, other):
        return self._binary(other, "sub_")

    def __mul__(self, other):
        return self._binary(other, "mul_")

    __radd__ = __add__
    __rmul__ = __mul__

    def __repr__(self):
        return "<NDArray %s @%s %s %s>" % ("x".join(map(str, self.shape)), self.context, self.dtype, self.stype)

    # -- row_sparse --------------------------------------------------------------
    @property
    def indices(self):
        out = ctypes.c_void_p()
        check_call(_LIB.MXNDArrayGetAuxNDArray(self.handle, 0, ctypes.byref(out)))
        return NDArray(out, keep=self)

    @property
    def data(self):
        out = ctypes.c_void_p()
        check_call(_LIB.MXNDArrayGetDataNDArray(self.handle, ctypes.byref(out)))
        return NDArray(out, keep=self)

    def todense_numpy(self):
        shape = self.shape
        idx = self.indices.asnumpy()
        t = self.mx_dtype
        out = np.zeros(shape, np.float32 if t == BFLOAT16 else _DTYPE_MX_TO_NP[t])
        if idx.size:
            out[idx] = self.data.asnumpy()
        return out


def empty(shape, ctx=None, dtype=np.float32, stype="default", capacity=None):
    """``capacity`` (row_sparse only): number of rows the array can hold (default: all rows)."""
    if isinstance(shape, int):
        shape = (shape,)
    ctx = ctx or cpu()
    cshape = (ctypes.c_int64 * len(shape))(*shape)
    out = ctypes.c_void_p()
    if stype == "row_sparse":
        aux_type = (ctypes.c_int * 1)(6)
        aux_ndims = (ctypes.c_int * 1)(1)
        aux_shape = (ctypes.c_int64 * 1)(shape[0] if capacity is None else max(int(capacity), 1))
        check_call(_LIB.MXNDArrayCreateSparseEx64(1, cshape, len(shape), ctx.device_typeid, ctx.device_id, 0,
                                                  _mx_dtype(dtype), 1, aux_type, aux_ndims, aux_shape,
                                                  ctypes.byref(out)))
    else:
        check_call(_LIB.MXNDArrayCreate64(cshape, len(shape), ctx.device_typeid, ctx.device_id, 0,
                                          _mx_dtype(dtype), ctypes.byref(out)))
    return NDArray(out)


def array(source, ctx=None, dtype=None):
    """ndarray.py:3378-3418: an NDArray source keeps its dtype; anything else becomes float32 unless ``dtype`` says
    otherwise (numpy's own dtype is NOT taken over)."""
    if isinstance(source, NDArray):
        if dtype is None:
            dtype = source.dtype
        src = source.asnumpy(raw=True) if source.dtype == "bfloat16" else source.asnumpy()
    else:
        src = np.asarray(source)
        if dtype is None:
            dtype = np.float32
    out = empty(src.shape, ctx, dtype)
    out._sync_copyfrom(src)
    return out


def zeros(shape, ctx=None, dtype=np.float32, stype="default"):
    if isinstance(shape, int):
        shape = (shape,)
    if stype == "row_sparse":
        return row_sparse_array(np.zeros(shape, np.float32), ctx=ctx, dtype=dtype)
    out = empty(shape, ctx, dtype)
    out._sync_copyfrom(np.zeros(shape, np.float32))
    return out


def ones(shape, ctx=None, dtype=np.float32):
    if isinstance(shape, int):
        shape = (shape,)
    out = empty(shape, ctx, dtype)
    out._sync_copyfrom(np.ones(shape, np.float32))
    return out


def full(shape, val, ctx=None, dtype=np.float32):
    if isinstance(shape, int):
        shape = (shape,)
    out = empty(shape, ctx, dtype)
    out._sync_copyfrom(np.full(shape, val, np.float32))
    return out


def row_sparse_array(arg, shape=None, ctx=None, dtype=np.float32, capacity=None):
    """row_sparse from a dense numpy array (non-zero rows kept) or a (data, indices) pair."""
    if isinstance(arg, tuple):
        data, indices = np.asarray(arg[0]), np.asarray(arg[1], np.int64)
        assert shape is not None
    else:
        dense = np.asarray(arg)
        shape = dense.shape
        flat = dense.reshape(shape[0], -1)
        indices = np.where(np.any(flat != 0, axis=1))[0].astype(np.int64)
        data = dense[indices]
    out = empty(shape, ctx, dtype, stype="row_sparse", capacity=capacity if capacity is not None else len(indices))
    vals = empty((len(indices),) + tuple(shape[1:]), cpu(), dtype)
    idx = empty((len(indices),), cpu(), np.int64)
    if len(indices):
        vals._sync_copyfrom(data)
        idx._sync_copyfrom(indices)
    check_call(_LIB.MXNDArraySyncCopyFromNDArray(out.handle, vals.handle, ctypes.c_int(-1)))
    check_call(_LIB.MXNDArraySyncCopyFromNDArray(out.handle, idx.handle, ctypes.c_int(0)))
    return out


def from_dlpack(obj):
    """Zero-copy import of a DLPack producer (an object with ``__dlpack__`` such as a torch tensor
    or an MXNet NDArray, or a raw ``dltensor`` PyCapsule) through MXNDArrayFromDLPack
    (include/mxnet/c_api.h:993).  The producer's memory is kept alive until the NDArray is freed."""
    cap = obj.__dlpack__() if hasattr(obj, "__dlpack__") else obj
    ctypes.pythonapi.PyCapsule_GetPointer.restype = ctypes.c_void_p
    ctypes.pythonapi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    ptr = ctypes.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
    out = ctypes.c_void_p()
    check_call(_LIB.MXNDArrayFromDLPack(ctypes.c_void_p(ptr), ctypes.c_bool(False), ctypes.byref(out)))
    # ownership of the DLManagedTensor moved to the NDArray: mark the capsule as consumed
    ctypes.pythonapi.PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]
    ctypes.pythonapi.PyCapsule_SetName(cap, b"used_dltensor")
    return NDArray(out, keep=cap)


_TORCH_TO_MX = None


def from_torch(t):
    """Borrow a contiguous torch tensor's memory (no copy).  The tensor is kept alive by the
    returned NDArray."""
    import torch
    global _TORCH_TO_MX
    if _TORCH_TO_MX is None:
        _TORCH_TO_MX = {torch.float32: 0, torch.float64: 1, torch.float16: 2, torch.uint8: 3, torch.int32: 4,
                        torch.int8: 5, torch.int64: 6, torch.bfloat16: 12}
    if not t.is_contiguous():
        raise MXNetError("from_torch needs a contiguous tensor")
    shape = tuple(t.shape) if t.dim() > 0 else (1,)
    cshape = (ctypes.c_int64 * len(shape))(*shape)
    if t.is_cuda:
        dev_type, dev_id = 2, t.device.index if t.device.index is not None else torch.cuda.current_device()
    else:
        dev_type, dev_id = (3 if t.is_pinned() else 1), 0
    out = ctypes.c_void_p()
    check_call(_LIB.MXKVB200NDArrayFromPtr(ctypes.c_void_p(t.data_ptr()), cshape, len(shape), dev_type, dev_id,
                                           _TORCH_TO_MX[t.dtype], ctypes.byref(out)))
    return NDArray(out, keep=t)


def empty_symmetric(shape, dtype=np.float32):
    """Collective allocation in the peer-mapped arena (one-process-per-GPU mode); every rank must
    call it in the same order with the same shape."""
    if isinstance(shape, int):
        shape = (shape,)
    cshape = (ctypes.c_int64 * len(shape))(*shape)
    out = ctypes.c_void_p()
    check_call(_LIB.MXKVB200NDArrayCreateSymmetric(cshape, len(shape), _mx_dtype(dtype), ctypes.byref(out)))
    return NDArray(out)


def empty_multicast(shape, dtype=np.float32, group=None):
    """Collective (one process per GPU): allocate through torch's symmetric memory -- plumbing that
    creates the CUDA multicast object on NVSwitch systems -- and wrap the peer pointers and the
    multicast alias.  Arrays created this way are exchanged by the NVLS kernel
    (multimem.ld_reduce / multimem.st); without multicast support they behave like
    ``empty_symmetric``.  Every rank must call it in the same order with the same shape."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    if isinstance(shape, int):
        shape = (shape,)
    mxdt = _mx_dtype(dtype)
    tdt = {0: torch.float32, 1: torch.float64, 2: torch.float16, 3: torch.uint8, 4: torch.int32, 5: torch.int8,
           6: torch.int64, BFLOAT16: torch.bfloat16}[mxdt]
    n = 1
    for d in shape:
        n *= d
    t = symm_mem.empty(max(n, 1), dtype=tdt, device="cuda")
    grp = group if group is not None else dist.group.WORLD
    h = symm_mem.rendezvous(t, grp.group_name)
    world = dist.get_world_size(grp)
    ptrs = (ctypes.c_void_p * world)(*[int(p) for p in h.buffer_ptrs])
    mc = int(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else 0
    cshape = (ctypes.c_int64 * len(shape))(*shape)
    out = ctypes.c_void_p()
    check_call(_LIB.MXKVB200NDArrayFromPeers(ptrs, world, ctypes.c_void_p(mc), cshape, len(shape), mxdt,
                                             ctypes.byref(out)))
    return NDArray(out, keep=(t, h))


def has_multicast(nd_array):
    """True if the array has an NVSwitch multicast alias (engine-owned arena memory, or memory wrapped with one)."""
    out = ctypes.c_int(0)
    check_call(_LIB.MXKVB200NDArrayHasMulticast(nd_array.handle, ctypes.byref(out)))
    return bool(out.value)


def multi_sum_sq(*arrays, **kwargs):
    """Sums of squares of several arrays in one launch (the reference's ``multi_sum_sq`` operator,
    src/operator/contrib/multi_sum_sq-inl.h:83-96): returns float32 [len(arrays)] on the arrays' GPU."""
    scale = float(kwargs.pop("scale", 1.0))
    out = kwargs.pop("out", None)
    assert not kwargs and arrays
    if out is None:
        out = empty((len(arrays),), ctx=arrays[0].context, dtype=np.float32)
    handles = (ctypes.c_void_p * len(arrays))(*[a.handle.value for a in arrays])
    check_call(_LIB.MXKVB200MultiSumSq(len(arrays), handles, ctypes.c_float(scale), out.handle))
    return out


def multi_all_finite(*arrays, **kwargs):
    """1.0 if every element of every array is finite, else 0.0 (``multi_all_finite``,
    src/operator/all_finite.cu:68-103).  ``init_output=False`` keeps a 0 already in ``out``."""
    init_output = bool(kwargs.pop("init_output", True))
    out = kwargs.pop("out", None)
    assert not kwargs and arrays
    if out is None:
        assert init_output
        out = empty((1,), ctx=arrays[0].context, dtype=np.float32)
    handles = (ctypes.c_void_p * len(arrays))(*[a.handle.value for a in arrays])
    check_call(_LIB.MXKVB200MultiAllFinite(len(arrays), handles, int(init_output), out.handle))
    return out


def waitall():
    check_call(_LIB.MXNDArrayWaitAll())


class _SparseNamespace(object):
    """``mx.nd.sparse``: the two constructors of python/mxnet/ndarray/sparse.py the KVStore tests use"""
    row_sparse_array = staticmethod(row_sparse_array)

    @staticmethod
    def zeros(stype, shape, ctx=None, dtype=np.float32):
        return zeros(shape, ctx, dtype, stype=stype)


sparse = _SparseNamespace()
