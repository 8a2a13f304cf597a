"""Dynamic loss scaling for the Trainer (python/mxnet/amp/loss_scaler.py:25-80, amp.py:290-298,374-400).

Only the part of AMP that touches the gradient path lives here: the loss scaler, its overflow check
(``multi_all_finite`` over the gradients, in chunks of 200 arrays like the reference) and the
``init_trainer`` / ``scale_loss`` / ``unscale`` helpers.  Operator casting lists are MXNet-operator
business and out of scope; with torch models use ``torch.autocast`` for that.

Two places can decide about an overflow:
* parameters updated outside the kvstore (``update_on_kvstore=False`` / ``kvstore=None``): the Trainer
  asks ``LossScaler.has_overflow(grads)`` before its local update and skips it, as gluon/trainer.py:445-448;
* parameters updated on the kvstore with LAMB / LANS / LARS created with ``skip_nonfinite=True``: the
  push itself leaves everything untouched on overflow and the Trainer feeds ``KVStore.overflow()`` to
  ``LossScaler.update`` (the reference cannot skip in that configuration at all).
"""
import contextlib
import logging

from . import ndarray as _nd


class LossScaler(object):
    """loss_scaler.py:25-80: scale 2^16, halved on overflow, doubled after 2000 clean steps, max 2^24."""

    def __init__(self, init_scale=2. ** 16, scale_seq_len=2000, max_loss_scale=2. ** 24):
        self._loss_scale = init_scale
        self._next_loss_scale = self._loss_scale
        self._max_loss_scale = max_loss_scale
        self._scale_seq_len = scale_seq_len
        self._unskipped = 0

    @property
    def loss_scale(self):
        return self._loss_scale

    def update(self, has_overflow):
        """The bookkeeping half of has_overflow (loss_scaler.py:67-79)."""
        self._loss_scale = self._next_loss_scale
        if has_overflow:
            self._next_loss_scale = self._loss_scale / 2.
            self._unskipped = 0
            logging.info("AMP: decreasing loss scale to %f", self._next_loss_scale)
        else:
            self._unskipped += 1
        if self._unskipped == self._scale_seq_len:
            self._unskipped = 0
            self._next_loss_scale = min(self._max_loss_scale, self._loss_scale * 2.)
            logging.info("AMP: increasing loss scale to %f", self._next_loss_scale)
        return has_overflow

    def has_overflow(self, grads):
        """grads: NDArrays on one GPU (the reference looks at ``p._grad[0]`` of every parameter)."""
        grads = [g for g in grads if g is not None]
        chunk = 200
        flag = _nd.ones((1,), ctx=grads[0].context)
        for i in range(0, len(grads), chunk):
            _nd.multi_all_finite(*grads[i:i + chunk], init_output=False, out=flag)
        return self.update(not bool(flag.asnumpy()[0]))


def init_trainer(trainer, loss_scaler=None):
    """amp.init_trainer (amp.py:374-400)."""
    trainer._amp_loss_scaler = loss_scaler or LossScaler()
    trainer._amp_original_scale = trainer._scale
    return trainer


@contextlib.contextmanager
def scale_loss(loss, trainer):
    """amp.scale_loss (amp.py:290-298): the optimizer's rescale_grad takes the scale back out."""
    assert getattr(trainer, "_amp_loss_scaler", None) is not None, \
        "Loss scaler is not initialized, did you forget to call amp.init_trainer()?"
    s = trainer._amp_loss_scaler.loss_scale
    trainer._scale = trainer._amp_original_scale / s
    if isinstance(loss, (list, tuple)):
        yield [l * s for l in loss]
    else:
        yield loss * s


def unscale(trainer):
    """amp.unscale (amp.py:402-417): multiply the gradients by the pending scale now."""
    trainer._bind_grads()
    for gs in trainer._grads:
        for g in gs:
            g *= trainer._scale
    trainer._scale = 1.
