"""Trainer: the driver loop of python/mxnet/gluon/trainer.py (step -> _allreduce_grads ->
kv.pushpull per parameter with priority -i, update on kvstore) for torch parameters.

``params`` is a list of torch Parameters -- or of any objects with ``.data`` / ``.grad`` holding torch tensors
or engine NDArrays -- (one replica per process: one-process-per-GPU mode or a single GPU), or a list of lists -- ``params[i][d]`` being parameter i's replica on device d --
for the reference's single-process multi-GPU layout (param.list_data()/list_grad()).
"""
from . import kvstore as _kv
from . import optimizer as _opt
from .ndarray import NDArray, from_torch


def _nd(x):
    """Engine view of a parameter's data or gradient: a torch tensor is wrapped zero-copy, an NDArray (the
    reference's own parameter type, and what the simulator tests use) is taken as it is."""
    return x if isinstance(x, NDArray) else from_torch(x)


def _ptr(x):
    return x.data_ptr if isinstance(x, NDArray) else x.data_ptr()


def _stype(x):
    return x.stype if isinstance(x, NDArray) else "default"


class Trainer(object):
    def __init__(self, params, optimizer, optimizer_params=None, kvstore="device", compression_params=None,
                 update_on_kvstore=None, batched=True, symmetric=False, overlap=False, bucket_bytes=32 << 20):
        if isinstance(params, dict):               # trainer.py:80-84: a dict is taken in the order of its keys
            params = [params[k] for k in sorted(params.keys())]
        if not isinstance(params, (list, tuple)):
            raise ValueError("First argument must be a list or dict of Parameters, got %s." % type(params))
        self._params = [p if isinstance(p, (list, tuple)) else [p] for p in params]
        optimizer_params = dict(optimizer_params or {})
        self._scale = float(optimizer_params.get("rescale_grad", 1.0))
        if isinstance(optimizer, _opt.Optimizer):
            assert not optimizer_params, "optimizer_params must be None if optimizer is an Optimizer instance"
            self._optimizer = optimizer
        else:
            self._optimizer = _opt.create(optimizer, **optimizer_params)
        self._attach_params()
        self._kvstore_arg = kvstore
        self._compression_params = compression_params
        self._update_on_kvstore_arg = update_on_kvstore
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._batched = batched
        self._symmetric = symmetric      # rebind p.data / p.grad to peer-mapped arena memory (zero-copy)
        # overlap: exchange + update a bucket of parameters as soon as backward has produced its gradients
        # (torch post-accumulate-grad hooks), on the engine's high-priority stream, while backward goes on -- the
        # role priority=-i plays on the reference's engine (trainer.py:386-409, threaded_engine_perdevice.cc:97-279)
        self._overlap = bool(overlap)
        self._bucket_bytes = int(bucket_bytes)
        self._buckets = None
        self._armed_batch = None
        self.trace = None                # set to [] to collect (bucket, gradients-ready event on the framework stream,
                                         # exchange-finished event on the engine stream)
        self._grads = None
        self._weights = None

    def _attach_params(self):
        """trainer.py:139-153,540-541: the optimizer sees the Parameter objects, whose ``lr_mult`` / ``wd_mult``
        it reads at every update (plain torch Parameters have neither and are left out)."""
        pd = {i: reps[0] for i, reps in enumerate(self._params)
              if hasattr(reps[0], "lr_mult") or hasattr(reps[0], "wd_mult")}
        if pd:
            self._optimizer.param_dict = pd

    def _reset_kvstore(self):
        """trainer.py:178-186: forget the store (and the optimizer state it holds); the next step creates it
        again and broadcasts the parameters as they are then -- what loading a checkpoint into the
        parameters triggers in the reference."""
        if self._kvstore is not None and "dist" in self._kvstore.type:
            raise RuntimeError("Cannot reset distributed KVStore.")
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._grads = None

    @property
    def optimizer(self):
        return self._optimizer

    @property
    def learning_rate(self):
        return self._optimizer.learning_rate

    def set_learning_rate(self, lr):
        self._optimizer.set_learning_rate(lr)

    def _init_kvstore(self):
        """trainer.py:188-277, single-machine rows of the decision table: row_sparse weights must be updated on
        the store; dense weights with row_sparse gradients default to per-device updates; dense / dense follows
        ``update_on_kvstore`` (default: MXNET_UPDATE_ON_KVSTORE, on; off for 'local' with a parameter above 16M
        elements, model.py:88-130) as far as the store is capable of it.  Unlike the reference a store is kept
        for a single replica per parameter as well: in one-process-per-GPU mode that replica is one of many, and
        the fused update inside the store is this engine's fast path either way.  For the same reason the
        reference's ``optimizer.aggregate_num > 1 -> update_on_kvstore=False`` rule (trainer.py:114-119) is not
        applied: it exists because only the per-device updaters can aggregate tensors there, while this store
        updates every key of a call in one launch (sequence)."""
        import os
        import numpy as np
        kv = self._kvstore_arg
        config_uok = self._update_on_kvstore_arg
        sparse_weight = any(_stype(reps[0].data) != "default" for reps in self._params)
        sparse_grad = any(getattr(reps[0], "grad", None) is not None and _stype(reps[0].grad) != "default"
                          for reps in self._params)
        if sparse_weight:
            if isinstance(kv, str):
                kv = _kv.create(kv)
            elif not isinstance(kv, _kv.KVStore):
                raise TypeError("Cannot create '%s' KVStore with row_sparse parameters. "
                                "The type must be KVStore or str." % kv)
            assert kv.is_capable(_kv.KVStoreBase.OPTIMIZER), "KVStore with sparse weight requires optimizer support."
            if config_uok is False:
                raise ValueError("Cannot set update_on_kvstore=False when sparse weights are present.")
            uok = True
        else:
            uok = bool(int(os.getenv("MXNET_UPDATE_ON_KVSTORE", "1")))
            if isinstance(kv, str):
                name = kv
                kv = _kv.create(kv)
                if name == "local" and max(int(np.prod(_nd(reps[0].data).shape)) for reps in self._params) > (16 << 20):
                    uok = False
            elif kv is not None and not isinstance(kv, _kv.KVStoreBase):
                raise TypeError("kvstore must be KVStore, str or None")
            if kv is None:
                uok = False
            else:
                uok = uok and bool(kv.is_capable(_kv.KVStoreBase.OPTIMIZER))
            distributed = kv is not None and "dist" in kv.type
            if sparse_grad:
                # one machine: per-device updates are usually faster; several machines: only the store can update,
                # because row_sparse_pull of a gradient does not exist (trainer.py:204-236)
                uok = distributed
                if config_uok is False and distributed:
                    raise ValueError("Cannot set update_on_kvstore=False on dist kvstore "
                                     "when sparse gradients are present.")
                if kv is not None and not isinstance(kv, _kv.KVStore):
                    raise ValueError("Cannot use {} for multi-device training with sparse gradients".format(type(kv)))
            if config_uok is not None and kv is not None:
                uok = config_uok
            if uok and not kv.is_capable(_kv.KVStoreBase.OPTIMIZER):
                if config_uok:
                    raise ValueError("Please set update_on_kvstore=False when training with " + str(type(kv)))
                uok = False
        self._kvstore = kv
        self._update_on_kvstore = uok
        # key lists in one call are this engine's extension of the plug-in API; a third-party store gets the
        # reference's call pattern, one key per call (trainer.py:155-176, 385-409)
        self._batched = self._batched and isinstance(kv, _kv.KVStore)
        if self._symmetric:
            self._bind_symmetric()
        else:
            self._weights = [[_nd(p.data) for p in reps] for reps in self._params]
        if kv is not None:
            if self._compression_params:
                kv.set_gradient_compression(self._compression_params)
            if uok:
                kv.set_optimizer(self._optimizer)
            # _init_params (trainer.py:155-176): broadcast(idx, w0, all_w); row_sparse weights are only stored
            dense = [i for i, w in enumerate(self._weights) if w[0].stype == "default"]
            for i, w in enumerate(self._weights):
                if w[0].stype != "default":
                    kv.init(i, w[0])
            if dense and isinstance(kv, _kv.KVStore):
                kv.broadcast(dense, [self._weights[i][0] for i in dense], [self._weights[i] for i in dense])
            else:
                for i in dense:
                    kv.broadcast(i, self._weights[i][0], self._weights[i])
        self._kv_initialized = True

    def _row_sparse_pull(self, parameter, out, row_id, full_idx=False):
        """trainer.py:310-324: the rows ``row_id`` of a row_sparse parameter (its index in ``params``, or the
        parameter object) into ``out``; all rows through a plain pull when ``full_idx``."""
        if not self._kv_initialized:
            self._init_kvstore()
        idx = parameter if isinstance(parameter, int) else \
            next(i for i, reps in enumerate(self._params) if parameter is reps[0] or parameter is reps)
        if full_idx:
            assert row_id.size == (out[0] if isinstance(out, (list, tuple)) else out).shape[0]
            self._kvstore.pull(idx, out=out, priority=-idx, ignore_sparse=False)
        else:
            self._kvstore.row_sparse_pull(idx, out=out, row_ids=row_id, priority=-idx)

    def _bind_symmetric(self):
        """One-process-per-GPU: move every parameter and its gradient into the engine's peer-mapped
        arena (collective allocation, same order on every rank) and hand torch views of that memory
        back to the module, so push/pull read and write them over NVLink without staging."""
        import torch
        from . import ndarray as _nd
        tmap = {torch.float32: "float32", torch.float16: "float16", torch.bfloat16: "bfloat16"}
        self._weights, self._grads, self._grad_ptrs = [], [], []
        for reps in self._params:
            assert len(reps) == 1, "symmetric binding is for one replica per process"
            p = reps[0]
            dt = tmap[p.dtype]
            shape = tuple(p.shape) if p.dim() > 0 else (1,)
            w = _nd.empty_symmetric(shape, dt if dt == "bfloat16" else getattr(__import__("numpy"), dt))
            g = _nd.empty_symmetric(shape, dt if dt == "bfloat16" else getattr(__import__("numpy"), dt))
            wt, gt = w.as_torch().view(p.shape), g.as_torch().view(p.shape)
            wt.copy_(p.data)
            gt.zero_()
            p.data = wt
            p.grad = gt
            self._weights.append([w])
            self._grads.append([g])
            self._grad_ptrs.append([gt.data_ptr()])
        torch.cuda.synchronize()

    def _bind_grads(self):
        self._grads = [[_nd(p.grad) for p in reps] for reps in self._params]
        self._grad_ptrs = [[_ptr(p.grad) for p in reps] for reps in self._params]
        self._sparse_idx = [i for i, g in enumerate(self._grads) if g[0].stype != "default"]
        held = set(self._sparse_idx)
        self._dense_idx = [i for i in range(len(self._params)) if i not in held]

    # -- overlap of the exchange with backward -----------------------------------------------------------------
    def _build_buckets(self):
        """Buckets in REVERSE parameter order (backward produces the last layers' gradients first), each about
        ``bucket_bytes`` of gradient; bucket b is one ``pushpull`` of its keys with priority -b."""
        import torch
        assert all(len(reps) == 1 and isinstance(reps[0].grad, torch.Tensor) or reps[0].grad is None
                   for reps in self._params), "overlap needs one torch replica per parameter"
        if self._grads is None:
            self._bind_grads()
        order = [i for i in reversed(range(len(self._params))) if self._params[i][0].requires_grad]
        buckets, cur, cur_bytes = [], [], 0
        for i in order:
            p = self._params[i][0]
            cur.append(i)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= self._bucket_bytes:
                buckets.append(sorted(cur)); cur, cur_bytes = [], 0
        if cur:
            buckets.append(sorted(cur))
        self._buckets = buckets
        self._bucket_of = {i: b for b, idx in enumerate(buckets) for i in idx}
        self._pending = [len(idx) for idx in buckets]
        self._fired = [False] * len(buckets)
        for i in order:
            self._params[i][0].register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(_param):
            if self._armed_batch is None:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._fire(b)
        return hook

    def _fire(self, b):
        idx = self._buckets[b]
        kv = self._kvstore
        ev = None
        if self.trace is not None:
            import torch
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        _kv.set_auto_fence(False)           # the framework stream must NOT wait for this exchange: backward goes on
        try:
            if ev is not None:
                ev[0].record()               # framework stream: the moment this bucket's gradients are complete
            kv.pushpull(idx, [self._grads[i] for i in idx], out=[self._weights[i] for i in idx], priority=-b)
            if ev is not None:
                ev[1].record(self._engine_stream())
                self.trace.append((b, ev[0], ev[1]))
        finally:
            _kv.set_auto_fence(True)
        self._fired[b] = True

    def _engine_stream(self):
        if getattr(self, "_estream", None) is None:
            import ctypes
            import torch
            from .base import _LIB, check_call
            dev = self._weights[0][0].context.device_id
            sp = ctypes.c_void_p()
            check_call(_LIB.MXKVB200GetEngineStream(dev, ctypes.byref(sp)))
            self._estream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", dev))
        return self._estream

    def arm(self, batch_size):
        """Overlap mode: the batch size of the coming ``step`` (rescale_grad is part of the fused update, which
        now starts during backward).  ``step`` arms the next iteration with its own batch size, so only a CHANGE
        of batch size has to be announced before ``backward``."""
        self._armed_batch = batch_size
        rescale = self._scale / batch_size
        if self._optimizer.rescale_grad != rescale or getattr(self._kvstore, "_last_rescale", None) != rescale:
            self._optimizer.rescale_grad = rescale
            if getattr(self._kvstore, "_fused", False):
                self._kvstore.set_optimizer(self._optimizer)
            self._kvstore._last_rescale = rescale

    def _finish_overlapped(self, batch_size):
        """the part of ``step`` that is left when the buckets were exchanged during backward"""
        assert self._armed_batch == batch_size, \
            "overlap mode: step(%r) after a backward armed for batch size %r; call trainer.arm(batch_size) before " \
            "backward when the batch size changes" % (batch_size, self._armed_batch)
        for b, fired in enumerate(self._fired):
            if not fired:                    # parameters that received no gradient in this backward
                self._fire(b)
        dev = self._weights[0][0].context.device_id
        _kv.fence(dev)                      # ONE stream wait: the framework stream sees every updated weight
        self._pending = [len(idx) for idx in self._buckets]
        self._fired = [False] * len(self._buckets)

    def _allreduce_grads(self):
        """trainer.py:385-409, with every parameter in ONE call when ``batched`` (one launch per
        GPU instead of one per parameter; the C ABI has always accepted key lists)."""
        if self._grads is None or any(_ptr(p.grad) != q for reps, ptrs in zip(self._params, self._grad_ptrs)
                                      for p, q in zip(reps, ptrs)):
            self._bind_grads()
        kv = self._kvstore
        idx = self._dense_idx if hasattr(self, "_dense_idx") else list(range(len(self._params)))
        sparse = getattr(self, "_sparse_idx", [])
        # row_sparse gradients: push and pull separately, parameter by parameter (trainer.py:392-401); the
        # pull brings back the weights when the store updates, the reduced gradient otherwise, and nothing for a
        # row_sparse weight (its rows are fetched on demand, _row_sparse_pull)
        for i in sparse:
            kv.push(i, self._grads[i], priority=-i)
            if self._weights[i][0].stype == "default":
                kv.pull(i, self._weights[i] if self._update_on_kvstore else self._grads[i], priority=-i,
                        ignore_sparse=False)
        if not idx:
            return
        if self._batched:
            grads = self._grads if not sparse else [self._grads[i] for i in idx]
            if self._update_on_kvstore:
                kv.pushpull(idx, grads, out=self._weights if not sparse else [self._weights[i] for i in idx],
                            priority=0)
            else:
                kv.pushpull(idx, grads, priority=0)
        else:
            # the reference's loop (trainer.py:386-409): one call per parameter, priority -i.  On this engine the
            # calls are queued and issued at the flush in priority order, merged into one launch (MXKVB200SetDeferred)
            native = isinstance(kv, _kv.KVStore)
            if native:
                kv.set_deferred(True)
            for i in idx:
                if self._update_on_kvstore:
                    kv.pushpull(i, self._grads[i], out=self._weights[i], priority=-i)
                else:
                    kv.pushpull(i, self._grads[i], priority=-i)
            if native:
                kv.flush()
                kv.set_deferred(False)

    def step(self, batch_size, ignore_stale_grad=False):
        """trainer.py:334-361: rescale_grad = scale / batch_size, allreduce, update."""
        if self._overlap and self._kv_initialized and self._buckets is not None and self._armed_batch is not None:
            self._finish_overlapped(batch_size)
            return
        self._optimizer.rescale_grad = self._scale / batch_size
        if not self._kv_initialized:
            self._init_kvstore()
        elif self._update_on_kvstore and self._kvstore is not None and \
                getattr(self._kvstore, "_fused", False) and \
                getattr(self._kvstore, "_last_rescale", None) != self._optimizer.rescale_grad:
            # only the fused engine holds a COPY of the hyper-parameters; a Python Updater reads
            # optimizer.rescale_grad live (updater.py:39-93), re-installing it would drop its states
            self._kvstore.set_optimizer(self._optimizer)
        scaler = getattr(self, "_amp_loss_scaler", None)
        if self._kvstore is not None:
            self._kvstore._last_rescale = self._optimizer.rescale_grad
            self._allreduce_grads()
            if scaler is not None and self._update_on_kvstore:
                # the push decided on the device (optimizer created with skip_nonfinite=True)
                assert getattr(self._optimizer, "skip_nonfinite", False), \
                    "AMP with update_on_kvstore needs an optimizer created with skip_nonfinite=True " \
                    "(lamb / lans / lars); otherwise use update_on_kvstore=False"
                scaler.update(self._kvstore.overflow())
        if not self._update_on_kvstore:
            self._update(ignore_stale_grad)
        if self._overlap and self._update_on_kvstore and self._batched and self._kvstore is not None and \
                getattr(self._kvstore, "_fused", False) and not getattr(self, "_sparse_idx", []) and scaler is None:
            # from the next backward on, buckets are exchanged as their gradients become ready
            if self._buckets is None:
                self._build_buckets()
            self.arm(batch_size)

    def allreduce_grads(self):
        if not self._kv_initialized:
            self._init_kvstore()
        assert not self._update_on_kvstore, \
            "allreduce_grads() when parameters are updated on kvstore is not supported."
        self._allreduce_grads()

    def update(self, batch_size, ignore_stale_grad=False):
        """trainer.py:411-442: the update half of ``step`` for callers that ran ``allreduce_grads()``
        themselves (e.g. to clip the reduced gradients); not available when the kvstore updates."""
        if not self._kv_initialized:
            self._init_kvstore()
        assert not (self._kvstore is not None and self._update_on_kvstore), \
            "update() when parameters are updated on kvstore is not supported. " \
            "Try setting `update_on_kvstore` to False when creating trainer."
        self._optimizer.rescale_grad = self._scale / batch_size
        self._update(ignore_stale_grad)

    def _update(self, ignore_stale_grad=False):
        """trainer.py:444-480 (gradient freshness is not tracked here, so ``ignore_stale_grad`` has nothing to do): per-device local updaters.  Optimizers with a fused kernel update every
        parameter of a device in one native launch (the reference's aggregated multi_* operators);
        the others run the generic Python updater per parameter."""
        if not hasattr(self, "_updaters"):
            ndev = len(self._params[0])
            self._updaters = [_opt.get_updater(self._optimizer) for _ in range(ndev)]
        if self._grads is None or any(_ptr(p.grad) != q for reps, ptrs in zip(self._params, self._grad_ptrs)
                                      for p, q in zip(reps, ptrs)):
            self._bind_grads()
        scaler = getattr(self, "_amp_loss_scaler", None)
        if scaler is not None and scaler.has_overflow([gs[0] for gs in self._grads]):
            return          # skip on overflow (trainer.py:445-448)
        idx = list(range(len(self._params)))
        for d, upd in enumerate(self._updaters):
            if isinstance(upd, _opt.NativeUpdater):
                upd(idx, [gs[d] for gs in self._grads], [ws[d] for ws in self._weights])
            else:
                for i in idx:
                    upd(i, self._grads[i][d], self._weights[i][d])

    def save_states(self, fname):
        assert self._kv_initialized
        if self._update_on_kvstore:
            self._kvstore.save_optimizer_states(fname, dump_optimizer=True)
        else:
            if not hasattr(self, "_updaters"):      # before the first step: the reference creates them at init
                self._updaters = [_opt.get_updater(self._optimizer) for _ in range(len(self._params[0]))]
            with open(fname, "wb") as f:
                f.write(self._updaters[0].get_states(dump_optimizer=True))

    def load_states(self, fname):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
            self._optimizer = self._kvstore._optimizer
        else:
            with open(fname, "rb") as f:
                blob = f.read()
            if not hasattr(self, "_updaters"):
                self._updaters = [_opt.get_updater(self._optimizer) for _ in range(len(self._params[0]))]
            for u in self._updaters:
                u.set_states(blob)
                u.optimizer = self._updaters[0].optimizer
            self._optimizer = self._updaters[0].optimizer
        self._attach_params()
