"""ctypes loader and error plumbing (python/mxnet/base.py:234-312 in the reference)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmxkv_b200.so")


class MXNetError(RuntimeError):
    """Error raised by the native library (reference: python/mxnet/error.py)."""


def _load():
    path = os.environ.get("MXKV_B200_LIBRARY_PATH", _LIB_PATH)
    if not os.path.exists(path):
        # build in-tree when a toolchain is present; otherwise fail loudly -- no fallback
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_mxkv_b200_build", os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as e:  # pylint: disable=broad-except
            raise MXNetError("native library %s is missing and could not be built: %s" % (path, e))
    lib = ctypes.CDLL(path, ctypes.RTLD_LOCAL)
    lib.MXGetLastError.restype = ctypes.c_char_p
    return lib


_LIB = _load()

NDArrayHandle = ctypes.c_void_p
KVStoreHandle = ctypes.c_void_p
mx_uint = ctypes.c_uint


def check_call(ret):
    """Raise MXNetError with the library's message when a C call returns non-zero."""
    if ret != 0:
        raise MXNetError(_LIB.MXGetLastError().decode("utf-8", "replace"))


def c_str(s):
    return ctypes.c_char_p(s.encode("utf-8"))


def c_str_array(strings):
    arr = (ctypes.c_char_p * len(strings))()
    arr[:] = [s.encode("utf-8") for s in strings]
    return arr


def c_array(ctype, values):
    return (ctype * len(values))(*values)


def c_handle_array(objs):
    arr = (ctypes.c_void_p * len(objs))()
    arr[:] = [o.handle for o in objs]
    return arr
