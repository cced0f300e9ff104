"""ctypes binding of libdcb200.so (the C ABI declared in include/dcb200.h).

There is no CPU fallback: if the library is missing or a call fails, this module
raises.  ``lib()`` loads lazily so that importing ``deepcalcium`` works on a box
without a GPU (the host logic and the symbol-export tests run there).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'libdcb200.so')

_lib = None

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_d = ctypes.c_double
c_p = ctypes.c_void_p
c_sz = ctypes.c_size_t

DCB_F32, DCB_BF16 = 0, 1
LOSS_IDS = {'binary_crossentropy': 0, 'weighted_binary_crossentropy': 1, 'dice_loss': 2, 'dicesq_loss': 3}


class DcbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DcbError('libdcb200.so not built (%s); run `python deep-calcium_b200/build.py` or '
                           '__graft_entry__.build() - there is no CPU fallback' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dcb_last_error.restype = ctypes.c_char_p
        _lib.dcb_launch_count.restype = ctypes.c_ulonglong
    return _lib


def check(status, what=''):
    if status != 0:
        msg = lib().dcb_last_error()
        raise DcbError('%s failed with status %d: %s' % (what or 'dcb call', status,
                                                        msg.decode() if msg else '?'))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return c_p(0)
    return c_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    fn = getattr(lib(), name)
    check(fn(*args), name)


def launch_count():
    return int(lib().dcb_launch_count())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise DcbError('deepcalcium (B200 build) needs a CUDA device; there is no CPU fallback')
    lib()
