"""Device contexts (python/mxnet/context.py; Context::DeviceType include/mxnet/base.h:94-99)."""
import ctypes
from .base import _LIB, check_call


class Context(object):
    devtype2str = {1: "cpu", 2: "gpu", 3: "cpu_pinned"}
    devstr2type = {"cpu": 1, "gpu": 2, "cpu_pinned": 3}

    def __init__(self, device_type, device_id=0):
        if isinstance(device_type, Context):
            self.device_typeid, self.device_id = device_type.device_typeid, device_type.device_id
        else:
            self.device_typeid = Context.devstr2type[device_type]
            self.device_id = device_id

    @property
    def device_type(self):
        return Context.devtype2str[self.device_typeid]

    def __eq__(self, other):
        return isinstance(other, Context) and self.device_typeid == other.device_typeid and \
            self.device_id == other.device_id

    def __hash__(self):
        return hash((self.device_typeid, self.device_id))

    def __repr__(self):
        return "%s(%d)" % (self.device_type, self.device_id)


def cpu(device_id=0):
    return Context("cpu", device_id)


def gpu(device_id=0):
    return Context("gpu", device_id)


def cpu_pinned(device_id=0):
    return Context("cpu_pinned", device_id)


def num_gpus():
    n = ctypes.c_int()
    check_call(_LIB.MXGetGPUCount(ctypes.byref(n)))
    return n.value
