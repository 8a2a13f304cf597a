"""``mx.gluon.Trainer`` -- the one gluon class on this path (python/mxnet/gluon/trainer.py); blocks, parameters and
data loading are not part of this package."""
from .trainer import Trainer  # noqa: F401
