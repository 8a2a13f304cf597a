"""The reduction trees of ``MXNET_KVSTORE_USETREE=1`` (reference: ``CommDeviceTree``, src/kvstore/comm_tree.h, and
its solver src/kvstore/gpu_topology.h) as the native library builds them -- ctypes views of the
``MXKVB200Topology*`` entry points, for tests and for a look at what a machine's trees are.

Nothing here is on the data path: a store created with ``MXNET_KVSTORE_USETREE=1`` builds its trees inside the
library (csrc/topology.cc) the first time a key is pushed from three or more GPUs."""
import ctypes

import numpy as np

from .base import _LIB, check_call


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def depth(n):
    """Levels below the root of a tree over n GPUs (ComputeDepth, gpu_topology.h:714-721)."""
    d = 1
    while n > (1 << d):
        d += 1
    return d


def link_weights(perf_rank, can_access):
    """GetP2PWeight (gpu_topology.h:137-253) applied to the driver's answers (n x n integer matrices)."""
    p = np.ascontiguousarray(perf_rank, np.int32)
    a = np.ascontiguousarray(can_access, np.int32)
    n = p.shape[0]
    out = np.empty((n, n), np.float32)
    check_call(_LIB.MXKVB200TopologyLinkWeights(n, _vp(p), _vp(a), _vp(out)))
    return out


def query_links(devs):
    """The link matrix of CUDA devices ``devs`` as the reference would derive it on this machine."""
    d = np.ascontiguousarray(devs, np.int32)
    out = np.empty((d.size, d.size), np.float32)
    check_call(_LIB.MXKVB200TopologyQueryLinks(int(d.size), _vp(d), _vp(out)))
    return out


def compute_trees(weights, alpha=0.7, backtrack=False):
    """ComputeTrees (gpu_topology.h:1111-1157): (topo [n, 2^(d+1)-1], scan [n, d+2], d)."""
    w = np.ascontiguousarray(weights, np.float32)
    n = w.shape[0]
    topo = np.zeros(n * (2 << depth(n)), np.uint64)
    scan = np.zeros(n * (depth(n) + 2), np.uint64)
    tl, sl, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check_call(_LIB.MXKVB200TopologyComputeTrees(_vp(w), n, ctypes.c_float(alpha), 1 if backtrack else 0, _vp(topo),
                                                 int(topo.size), ctypes.byref(tl), _vp(scan), int(scan.size),
                                                 ctypes.byref(sl), ctypes.byref(d)))
    return (topo[:n * tl.value].reshape(n, tl.value).astype(np.int64),
            scan[:n * sl.value].reshape(n, sl.value).astype(np.int64), d.value)


def bisect(weights, partition, num_partitions, seed=1):
    """One Kernighan-Lin pass (gpu_topology.h:326-480): (stop, partition, num_partitions, pairs)."""
    w = np.ascontiguousarray(weights, np.float32)
    n = w.shape[0]
    p = np.ascontiguousarray(partition, np.int32).copy()
    npart = ctypes.c_int(num_partitions)
    pairs = np.zeros(4 * n + 4, np.int32)
    npairs, stop = ctypes.c_int(), ctypes.c_int()
    check_call(_LIB.MXKVB200TopologyBisect(_vp(w), n, _vp(p), ctypes.byref(npart), _vp(pairs), int(pairs.size // 2),
                                           ctypes.byref(npairs), ctypes.c_uint32(seed), ctypes.byref(stop)))
    return bool(stop.value), p, npart.value, [tuple(pairs[2 * i:2 * i + 2]) for i in range(npairs.value)]


def fold_repeats(leaves, n, d):
    """Postprocess (gpu_topology.h:746-770)."""
    r = np.ascontiguousarray(leaves, np.int32).copy()
    check_call(_LIB.MXKVB200TopologyFoldRepeats(_vp(r), int(r.size), n, d))
    return r


def tree_weight(weights, leaves, n, d, penalty):
    """ComputeTreeWeight (gpu_topology.h:778-813)."""
    w = np.ascontiguousarray(weights, np.float32)
    r = np.ascontiguousarray(leaves, np.int32)
    out = ctypes.c_float()
    check_call(_LIB.MXKVB200TopologyTreeWeight(_vp(w), _vp(r), int(r.size), n, d, 1 if penalty else 0, ctypes.byref(out)))
    return out.value


def admissible(weights, state, n, row, d):
    """IsValid (gpu_topology.h:727-791)."""
    w = np.ascontiguousarray(weights, np.float32)
    s = np.ascontiguousarray(state, np.int32)
    out = ctypes.c_int()
    check_call(_LIB.MXKVB200TopologyAdmissible(_vp(w), _vp(s), int(s.size), n, row, d, ctypes.byref(out)))
    return bool(out.value)


def connected(weights):
    """IsConnected (gpu_topology.h:96-121)."""
    w = np.ascontiguousarray(weights, np.float32)
    out = ctypes.c_int()
    check_call(_LIB.MXKVB200TopologyConnected(_vp(w), int(w.shape[0]), ctypes.byref(out)))
    return bool(out.value)


def reduce_program(topo_row, scan_row, n):
    """(leaf order, add schedule) of one tree: the order in which the tree kernel takes the n values of an element
    and the bit string that says when partial sums meet (csrc/topology.h: ReduceProgram)."""
    t = np.ascontiguousarray(topo_row, np.uint64)
    s = np.ascontiguousarray(scan_row, np.uint64)
    leaves = np.zeros(n, np.int32)
    prog = ctypes.c_uint32()
    check_call(_LIB.MXKVB200TopologyReduceProgram(_vp(t), int(t.size), _vp(s), int(s.size), n, _vp(leaves),
                                                  ctypes.byref(prog)))
    return leaves, prog.value


def run_program(srcs_in_leaf_order, prog):
    """The kernel's evaluator (csrc/tree_math.h) run on the host over float32 arrays."""
    arrs = [np.ascontiguousarray(a, np.float32) for a in srcs_in_leaf_order]
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.empty_like(arrs[0])
    check_call(_LIB.MXKVB200TopologyRunProgram(ptrs, len(arrs), ctypes.c_uint32(prog), ctypes.c_int64(arrs[0].size), _vp(out)))
    return out
