"""Regenerates tests/golden/*.npz.  Runs ONLY in the authoring container (needs /root/reference):

* dense_sums.npz   -- outputs of the reference's own mshadow arithmetic (oracle/_ref/libkvref.so,
                      built from /root/reference/3rdparty/mshadow by oracle/Makefile) for the
                      CommDevice and CommCPU reduce orders;
* compression.npz  -- outputs of the reference's bit-level simulator `compute_1bit` /
                      `compute_2bit` (tests/nightly/test_kvstore.py:35-98), executed from the
                      reference file itself (nothing is copied into this repository).

The GPU box has no /root/reference; tests read the committed fixtures instead.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def dense_sums():
    from oracle import oracle as O
    assert O.ref_lib() is not None, "build oracle/_ref first (make -C oracle)"
    out = {}
    rng = np.random.default_rng(20260921)
    for n in (2, 3, 4, 5, 7, 8):
        for E in (1, 33, 4099):
            vals = [rng.uniform(-1, 1, E).astype(np.float32) for _ in range(n)]
            tag = "n%d_E%d" % (n, E)
            out["in_" + tag] = np.stack(vals)
            out["device_f32_" + tag] = O.ref_sum_device(vals)
            out["commcpu_f32_" + tag] = O.ref_sum_cpu(vals)
            out["device_f16_" + tag] = O.ref_sum_device([v.astype(np.float16) for v in vals])
    # big enough to take CommCPU's OpenMP branch (>= 1e6 elements, 4096-element tasks)
    vals = [rng.uniform(-1, 1, 1000003).astype(np.float32) for _ in range(5)]
    out["seed_big"] = np.array([777], np.int64)
    big_rng = np.random.default_rng(777)
    vals = [big_rng.uniform(-1, 1, 1000003).astype(np.float32) for _ in range(5)]
    s = O.ref_sum_cpu(vals)
    out["commcpu_f32_big_head"] = s[:4096]
    out["commcpu_f32_big_tail"] = s[-4096:]
    out["commcpu_f32_big_xor"] = np.bitwise_xor.reduce(s.view(np.uint32)).reshape(1)
    np.savez_compressed(os.path.join(HERE, "dense_sums.npz"), **out)


def compression():
    src = open(os.path.join(REF, "tests/nightly/test_kvstore.py")).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("compute_1bit", "compute_2bit"):
            exec(compile(ast.Module([node], []), "reference:tests/nightly/test_kvstore.py", "exec"), ns)

    def pack(bits):
        # compute_expected_quantization (:81-98) reverses the four bytes of every 32-bit group and
        # stores the word little-endian, i.e. memory byte j holds bits[8j:8j+8] MSB first
        by = [int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)]
        return np.array(by, np.uint8)

    out = {}
    rng = np.random.default_rng(4242)
    for kind, fn, thr in (("2bit", ns["compute_2bit"], 0.5), ("1bit", ns["compute_1bit"], 0.0)):
        for E in (32, 64, 160):
            res = np.zeros(E, np.float32)
            for it in range(3):
                arr = rng.uniform(-1.2, 1.2, E).astype(np.float32)
                bits, new_res, dec = fn(arr.copy(), res, thr)
                tag = "%s_E%d_it%d" % (kind, E, it)
                out["grad_" + tag] = arr
                out["res_in_" + tag] = res.copy()
                out["bytes_" + tag] = pack(bits)[: ((E + (31 if kind == "1bit" else 15)) // (32 if kind == "1bit" else 16)) * 4]
                out["res_out_" + tag] = np.array(new_res, np.float32)
                out["dec_" + tag] = np.array(dec, np.float32)
                res = np.array(new_res, np.float32)
    np.savez_compressed(os.path.join(HERE, "compression.npz"), **out)


if __name__ == "__main__":
    dense_sums()
    compression()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
