"""Worker of test_sim_host_logic.py::test_layerwise_kernels_from_source_match_the_emulators_bit_for_bit: a few LAMB /
LANS / LARS steps on 1 ... 3 simulated GPUs; prints a digest of every result.  Run once with the kernels' own source
(512-thread blocks as user-level contexts, warp shuffles as lane exchanges) and once with MXKV_SIM_NORM=semantic
(the emulators of sim_kernels.cc, which restate the hardware's summation order): the digests must be equal."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxnet_b200 as mx  # noqa: E402

rng = np.random.default_rng(1)
shapes = [(1000,), (64, 33), (200003,), (8192 * 3 + 5,)]
keys = list(range(len(shapes)))
h = hashlib.sha256()
for name, kw in (("LAMB", dict(learning_rate=0.01, wd=0.01)), ("LANS", dict(learning_rate=0.01, wd=0.01)),
                 ("LARS", dict(learning_rate=0.1, momentum=0.9, wd=1e-3, eta=0.01))):
    for ndev in (1, 2, 3):
        kv = mx.kv.create("device")
        w0 = [rng.uniform(0, 1, s).astype(np.float32) for s in shapes]
        kv.init(keys, [mx.nd.array(w, mx.gpu(0)) for w in w0])
        kv.set_optimizer(getattr(mx.optimizer, name)(**kw))
        outs = [mx.nd.empty(s, mx.gpu(0)) for s in shapes]
        for step in range(3):
            grads = [[rng.uniform(-1, 1, s).astype(np.float32) for _ in range(ndev)] for s in shapes]
            kv.pushpull(keys, [[mx.nd.array(g, mx.gpu(d)) for d, g in enumerate(gs)] for gs in grads], out=outs)
        for o in outs:
            h.update(o.asnumpy().tobytes())
print("DIGEST", h.hexdigest())
