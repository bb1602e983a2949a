"""ctypes loader for libdiffsal_b200.so.  There is no fallback: if the library is missing or
no CUDA device is present the product path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffsal_b200.so")
_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                "%s not built: run `python -m diff_sal_b200.build` (or __graft_entry__.build()). "
                "There is no CPU / PyTorch fallback for the denoiser." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
