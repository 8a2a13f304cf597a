"""The arithmetic source of the CUDA kernels (csrc/optim_math.h, csrc/norm_math.h) compiled for the HOST
and checked bit-for-bit against the oracle: parity of the optimizer arithmetic that needs no GPU.

The kernels spell every operation with __fmul_rn / __fadd_rn / ... so that ptxas cannot contract them; on the
host those names are plain IEEE single-precision operations (tests/c/optim_host.cc, built with
-ffp-contract=off), which is exactly what the device intrinsics compute.  The oracle is the reference's
arithmetic restated in C (oracle/kv_oracle.c), pinned to the reference's own Python `step` outputs
(tests/golden)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32P = ctypes.POINTER(ctypes.c_float)
OPT = dict(none=0, sgd=1, sgd_mom=2, adam=3, adamw=4, test=5, sgd_std=6, adam_std=7)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no C++ compiler")
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    so = str(tmp_path_factory.mktemp("optim_host") / "liboptim_host.so")
    subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Werror", "-Wno-unknown-pragmas",
                    "-I", os.path.join(ROOT, "incubator-mxnet_b200", "csrc"), "-I", cuda_inc,
                    os.path.join(ROOT, "tests", "c", "optim_host.cc"), "-o", so], check=True, capture_output=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(F32P) if a is not None else None


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _update(host, opt, g, w, s0=None, s1=None, lr=0.0, wd=0.0, eta=1.0, rescale=1.0, clip=None, momentum=0.0,
            beta1=0.9, beta2=0.999, eps=1e-8):
    f = ctypes.c_float
    rc = host.host_update(OPT[opt], ctypes.c_int64(w.size), _p(g), _p(w), _p(s0), _p(s1), f(lr), f(wd), f(eta),
                          f(rescale), f(-1.0 if clip is None else clip), f(momentum), f(beta1), f(beta2), f(eps))
    assert rc == 0


def _data(seed, n=20011):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32))


@pytest.mark.parametrize("clip", [None, 0.3])
def test_sgd_and_momentum(host, clip):
    g, w = _data(1)
    kw = dict(lr=0.1, wd=1e-3, rescale=0.5, clip=clip)
    w1, w2 = w.copy(), w.copy()
    _update(host, "sgd", g, w1, **kw)
    O.sgd_update(w2, g, 0.1, 1e-3, 0.5, clip)
    assert _bits_equal(w1, w2)
    w1, w2, m1, m2 = w.copy(), w.copy(), np.zeros_like(w), np.zeros_like(w)
    for _ in range(3):
        _update(host, "sgd_mom", g, w1, m1, momentum=0.9, **kw)
        O.sgd_mom_update(w2, g, m2, 0.1, 1e-3, 0.9, 0.5, clip)
    assert _bits_equal(w1, w2) and _bits_equal(m1, m2)


def test_adam_adamw_test(host):
    g, w = _data(2)
    w1, w2 = w.copy(), w.copy()
    m1, v1, m2, v2 = (np.zeros_like(w) for _ in range(4))
    for t in range(1, 4):
        lr = np.float32(O.adam_lr(0.01, 0.9, 0.999, t))
        _update(host, "adam", g, w1, m1, v1, lr=lr, wd=1e-2, rescale=0.25, clip=0.8, eps=1e-8)
        O.adam_update(w2, g, m2, v2, lr, 1e-2, 0.9, 0.999, 1e-8, 0.25, 0.8)
    assert _bits_equal(w1, w2) and _bits_equal(m1, m2) and _bits_equal(v1, v2)
    # AdamW the way the reference's optimizer class drives the operator: lr = 1, eta = lr_t
    w1, w2 = w.copy(), w.copy()
    m1, v1, m2, v2 = (np.zeros_like(w) for _ in range(4))
    for t in range(1, 4):
        eta = np.float32(O.adam_lr(0.01, 0.9, 0.98, t))
        _update(host, "adamw", g, w1, m1, v1, lr=1.0, eta=eta, wd=0.05, beta2=0.98, eps=1e-6, clip=0.5)
        O.mp_adamw_update(None, 0, w2, m2, v2, g, 1.0, float(eta), 0.05, 0.9, 0.98, 1e-6, 1.0, 0.5)
    assert _bits_equal(w1, w2) and _bits_equal(m1, m2) and _bits_equal(v1, v2)
    w1, w2 = w.copy(), w.copy()
    _update(host, "test", g, w1, lr=0.3, wd=1e-2, rescale=0.5)
    O.test_update(w2, g, 0.3, 1e-2, 0.5)
    assert _bits_equal(w1, w2)


def test_standard_sparse_flavours(host):
    """OPT_SGD_STD / OPT_ADAM_STD on the densified gradient == the reference's *Std* sparse kernels."""
    rows, L = 400, 16
    rng = np.random.default_rng(3)
    w = rng.uniform(0, 1, (rows, L)).astype(np.float32)
    idx = np.sort(rng.choice(rows, 90, replace=False)).astype(np.int64)
    val = rng.uniform(-1, 1, (90, L)).astype(np.float32)
    rsp = O.RowSparse(idx, val, (rows, L))
    dense = np.ascontiguousarray(rsp.todense())
    w1, w2 = w.copy(), w.copy()
    _update(host, "sgd_std", dense.ravel(), w1.reshape(-1), lr=0.1, wd=1e-2, rescale=0.5, clip=0.4)
    O.sgd_std_rsp(w2, rsp, 0.1, 1e-2, 0.5, 0.4)
    assert _bits_equal(w1, w2)
    w1, w2 = w.copy(), w.copy()
    m1, v1, m2, v2 = (np.zeros_like(w) for _ in range(4))
    _update(host, "adam_std", dense.ravel(), w1.reshape(-1), m1.reshape(-1), v1.reshape(-1), lr=0.01, wd=1e-2,
            rescale=0.5, clip=0.4)
    O.adam_std_update(w2, dense, m2, v2, 0.01, 1e-2, 0.9, 0.999, 1e-8, 0.5, 0.4)
    assert _bits_equal(w1, w2) and _bits_equal(m1, m2) and _bits_equal(v1, v2)


@pytest.mark.parametrize("bias_correction", [True, False])
def test_lamb_steps(host, bias_correction):
    g, w = _data(4, 5003)
    f = ctypes.c_float
    mean1, var1, mean2, var2 = (np.zeros_like(w) for _ in range(4))
    w1, w2 = w.copy(), w.copy()
    lib = O.lib()
    for t in range(1, 4):
        c1 = np.float32(1.0) - np.float32(np.float32(0.9) ** np.float32(t))
        c2 = np.float32(1.0) - np.float32(np.float32(0.999) ** np.float32(t))
        gh1, gh2 = np.empty_like(w), np.empty_like(w)
        host.host_lamb_step1(ctypes.c_int64(w.size), _p(g), _p(w1), _p(mean1), _p(var1), _p(gh1), f(0.03), f(c1),
                             f(c2), f(0.5), f(0.4), f(0.9), f(0.999), f(1e-6), int(bias_correction))
        lib.kvo_lamb_step1_f32(ctypes.c_int64(w.size), _p(w2), _p(g), _p(mean2), _p(var2), _p(gh2), f(0.9), f(0.999),
                               f(1e-6), f(0.03), f(0.5), f(0.4), int(bias_correction), t)
        assert _bits_equal(gh1, gh2) and _bits_equal(mean1, mean2) and _bits_equal(var1, var2), t
        # step 2: the per-key scalar from the same totals, then w -= sc * ghat (kv_norm_apply_kernel)
        ssw, ssg = O.sum_sq(w1, mode="f64"), O.sum_sq(gh1, mode="f64")
        totals = np.array([ssw, ssg, 0, 0, 0, 0, 0, 0], np.float32)
        sc = np.zeros(2, np.float32)
        host.host_apply_scalars(0, _p(totals), f(0.01), ctypes.c_double(0.01), f(0.03), f(1e-3), f(10.0), f(0.9),
                                f(0.0), f(0.0), 0, _p(sc))
        w1 = (w1 - sc[0] * gh1).astype(np.float32)
        lib.kvo_lamb_step2_f32(ctypes.c_int64(w.size), _p(w2), _p(gh2), f(0.01), f(ssw), f(ssg), f(1e-3), f(10.0))
        assert _bits_equal(w1, w2), t


def test_lans_steps(host):
    g, w = _data(5, 5003)
    f = ctypes.c_float
    lib = O.lib()
    mean1, var1, mean2, var2 = (np.zeros_like(w) for _ in range(4))
    w1, w2 = w.copy(), w.copy()
    for t in range(1, 4):
        c1 = np.float32(1.0) - np.float32(np.float32(0.9) ** np.float32(t))
        c2 = np.float32(1.0) - np.float32(np.float32(0.999) ** np.float32(t))
        gsq = O.sum_sq(g, 0.5, mode="f64")
        tm1, tg1, tm2, tg2 = (np.empty_like(w) for _ in range(4))
        host.host_lans_step1(ctypes.c_int64(w.size), _p(g), _p(w1), _p(mean1), _p(var1), _p(tm1), _p(tg1), f(gsq),
                             f(0.03), f(c1), f(c2), f(0.5), f(0.4), f(0.9), f(0.999), f(1e-6))
        lib.kvo_lans_step1_f32(ctypes.c_int64(w.size), _p(w2), _p(g), _p(mean2), _p(var2), _p(tm2), _p(tg2), f(0.9),
                               f(0.999), f(1e-6), f(0.03), f(0.5), f(0.4), t, f(gsq))
        assert _bits_equal(tm1, tm2) and _bits_equal(tg1, tg2) and _bits_equal(mean1, mean2) and _bits_equal(var1, var2)
        ssw, ssm, ssg = O.sum_sq(w1, mode="f64"), O.sum_sq(tm1, mode="f64"), O.sum_sq(tg1, mode="f64")
        totals = np.array([ssw, 0, ssm, ssg, 0, 0, 0, 0], np.float32)
        sc = np.zeros(2, np.float32)
        host.host_apply_scalars(1, _p(totals), f(0.01), ctypes.c_double(0.01), f(0.03), f(-1.0), f(-1.0), f(0.9),
                                f(0.0), f(0.0), 0, _p(sc))
        w1 = (w1 - (sc[0] * tm1 + sc[1] * tg1)).astype(np.float32)
        lib.kvo_lans_step2_f32(ctypes.c_int64(w.size), _p(w2), _p(tm2), _p(tg2), f(0.01), f(0.9), f(ssw), f(ssm),
                               f(ssg), f(-1.0), f(-1.0))
        assert _bits_equal(w1, w2), t


def test_lars_ratio_and_trust_rules(host):
    f = ctypes.c_float
    lib = O.lib()
    rng = np.random.default_rng(6)
    for _ in range(200):
        ssw, ssg = np.float32(rng.uniform(0, 50)), np.float32(rng.uniform(0, 50))
        wd, lr = float(rng.uniform(0, 0.1)), float(rng.uniform(0.001, 1.0))
        totals = np.array([ssw, ssg, 0, 0, 0, 0, 0, 0], np.float32)
        sc = np.zeros(2, np.float32)
        host.host_apply_scalars(2, _p(totals), f(lr), ctypes.c_double(lr), f(wd), f(-1.0), f(-1.0), f(0.9), f(0.01),
                                f(1e-8), 0, _p(sc))
        lars = lib.kvo_lars_ratio_f32(f(ssw), f(ssg), f(0.01), f(wd), f(1e-8))
        assert sc[0] == np.float32(lr * float(lars))
    # zero weight norm, zero gradient norm, non-finite: ratio 1 (lars.py:127-131); no-trust flag: plain lr
    for ssw, ssg in ((0.0, 4.0), (4.0, 0.0), (np.inf, 1.0), (np.nan, 1.0)):
        totals = np.array([ssw, ssg, 0, 0, 0, 0, 0, 0], np.float32)
        sc = np.zeros(2, np.float32)
        host.host_apply_scalars(2, _p(totals), f(0.1), ctypes.c_double(0.1), f(0.01), f(-1.0), f(-1.0), f(0.9),
                                f(0.01), f(1e-8), 0, _p(sc))
        assert sc[0] == np.float32(0.1), (ssw, ssg, sc[0])
    totals = np.array([9.0, 4.0, 0, 0, 0, 0, 0, 0], np.float32)
    sc = np.zeros(2, np.float32)
    host.host_apply_scalars(2, _p(totals), f(0.1), ctypes.c_double(0.1), f(0.01), f(-1.0), f(-1.0), f(0.9), f(0.01),
                            f(1e-8), 2, _p(sc))
    assert sc[0] == np.float32(0.1)
    # LAMB: zero norms -> ratio 1 (multi_lamb.cc:104-107)
    for tot in ((0.0, 4.0), (4.0, 0.0)):
        totals = np.array([tot[0], tot[1], 0, 0, 0, 0, 0, 0], np.float32)
        host.host_apply_scalars(0, _p(totals), f(0.2), ctypes.c_double(0.2), f(0.0), f(-1.0), f(-1.0), f(0.9), f(0.0),
                                f(0.0), 0, _p(sc))
        assert sc[0] == np.float32(0.2)
    assert host.host_not_finite(f(np.inf)) == 1 and host.host_not_finite(f(np.nan)) == 1
    assert host.host_not_finite(f(-np.inf)) == 1 and host.host_not_finite(f(3.0e38)) == 0
