"""Learning-rate schedules an ``Optimizer`` consults with its ``num_update`` (python/mxnet/lr_scheduler.py):
``Factor``, ``MultiFactor``, ``Poly`` and ``Cosine`` decay, each with an optional warm-up.

The store reads ``optimizer.learning_rate`` right after counting the update and hands it to the engine
(``KVStore._advance_counts``), so a schedule behaves exactly as it does inside the reference's updater.
Values are computed with the same double arithmetic, in the same order, as the reference classes (the
committed fixture ``tests/golden/lr_schedules.npz`` holds the reference module's own outputs).
"""
import math


class LRScheduler(object):
    """Base: ``base_lr`` plus warm-up from ``warmup_begin_lr`` over ``warmup_steps`` updates
    (lr_scheduler.py:22-83)."""

    def __init__(self, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0, warmup_mode="linear"):
        if not isinstance(warmup_steps, int) or warmup_steps < 0:
            raise ValueError("warmup_steps must be a non-negative integer")
        if warmup_begin_lr > base_lr:
            raise ValueError("warmup_begin_lr may not exceed base_lr")
        if warmup_mode not in ("linear", "constant"):
            raise ValueError("warmup_mode is 'linear' or 'constant'")
        self.base_lr = base_lr
        self.warmup_steps = warmup_steps
        self.warmup_begin_lr = warmup_begin_lr
        self.warmup_final_lr = base_lr
        self.warmup_mode = warmup_mode

    def get_warmup_lr(self, num_update):
        assert num_update < self.warmup_steps
        if self.warmup_mode == "constant":
            return self.warmup_begin_lr
        span = self.warmup_final_lr - self.warmup_begin_lr
        return self.warmup_begin_lr + span * float(num_update) / float(self.warmup_steps)

    def _in_warmup(self, num_update):
        return num_update < self.warmup_steps

    def __call__(self, num_update):
        raise NotImplementedError("must override this")


class FactorScheduler(LRScheduler):
    """``base_lr * factor^floor(num_update / step)``, floored at ``stop_factor_lr`` (lr_scheduler.py:86-128).
    Stateful like the reference: the decay is applied incrementally, so resuming at a later ``num_update``
    catches up."""

    def __init__(self, step, factor=1, stop_factor_lr=1e-8, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0,
                 warmup_mode="linear"):
        super(FactorScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        if step < 1:
            raise ValueError("step must be at least 1")
        if factor > 1.0:
            raise ValueError("factor must not exceed 1")
        self.step, self.factor, self.stop_factor_lr = step, factor, stop_factor_lr
        self.count = 0

    def __call__(self, num_update):
        if self._in_warmup(num_update):
            return self.get_warmup_lr(num_update)
        while num_update > self.count + self.step:
            self.count += self.step
            self.base_lr = max(self.base_lr * self.factor, self.stop_factor_lr)
        return self.base_lr


class MultiFactorScheduler(LRScheduler):
    """Multiply by ``factor`` each time ``num_update`` passes the next entry of the increasing list ``step``
    (lr_scheduler.py:131-187)."""

    def __init__(self, step, factor=1, base_lr=0.01, warmup_steps=0, warmup_begin_lr=0, warmup_mode="linear"):
        super(MultiFactorScheduler, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        if not (isinstance(step, list) and step):
            raise ValueError("step must be a non-empty list")
        if any(s < 1 for s in step) or any(b <= a for a, b in zip(step, step[1:])):
            raise ValueError("step must be an increasing list of integers >= 1")
        if factor > 1.0:
            raise ValueError("factor must not exceed 1")
        self.step, self.factor = step, factor
        self.cur_step_ind = 0
        self.count = 0

    def __call__(self, num_update):
        if self._in_warmup(num_update):
            return self.get_warmup_lr(num_update)
        while self.cur_step_ind < len(self.step) and num_update > self.step[self.cur_step_ind]:
            self.count = self.step[self.cur_step_ind]
            self.cur_step_ind += 1
            self.base_lr *= self.factor
        return self.base_lr


class _ToFinal(LRScheduler):
    """Shared part of the schedules that travel from ``base_lr`` to ``final_lr`` by ``max_update``."""

    def __init__(self, max_update, base_lr, final_lr, warmup_steps, warmup_begin_lr, warmup_mode):
        super(_ToFinal, self).__init__(base_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        if not isinstance(max_update, int) or max_update < 1:
            raise ValueError("max_update must be a positive integer")
        self.base_lr_orig = self.base_lr
        self.max_update = max_update
        self.final_lr = final_lr
        self.max_steps = self.max_update - self.warmup_steps

    def _shape(self, num_update):
        raise NotImplementedError()

    def __call__(self, num_update):
        if self._in_warmup(num_update):
            return self.get_warmup_lr(num_update)
        if num_update <= self.max_update:
            self.base_lr = self.final_lr + (self.base_lr_orig - self.final_lr) * self._shape(num_update)
        return self.base_lr


class PolyScheduler(_ToFinal):
    """``final + (base - final) * (1 - t / T)^pwr`` (lr_scheduler.py:190-235)."""

    def __init__(self, max_update, base_lr=0.01, pwr=2, final_lr=0, warmup_steps=0, warmup_begin_lr=0,
                 warmup_mode="linear"):
        super(PolyScheduler, self).__init__(max_update, base_lr, final_lr, warmup_steps, warmup_begin_lr, warmup_mode)
        self.power = pwr

    def _shape(self, num_update):
        return pow(1 - float(num_update - self.warmup_steps) / float(self.max_steps), self.power)


class CosineScheduler(_ToFinal):
    """``final + (base - final) * (1 + cos(pi * t / T)) / 2`` (lr_scheduler.py:238-281)."""

    def __init__(self, max_update, base_lr=0.01, final_lr=0, warmup_steps=0, warmup_begin_lr=0,
                 warmup_mode="linear"):
        super(CosineScheduler, self).__init__(max_update, base_lr, final_lr, warmup_steps, warmup_begin_lr,
                                              warmup_mode)

    def _shape(self, num_update):
        return (1 + math.cos(math.pi * (num_update - self.warmup_steps) / self.max_steps)) / 2
