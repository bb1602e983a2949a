"""ctypes loader for libdiffsal_b200.so.  There is no fallback: if the library is missing or
no CUDA device is present the product path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffsal_b200.so")
TEST_LIB_PATH = os.path.join(_HERE, "libdiffsal_b200_test.so")
_lib = None
_test_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                "%s not built: run `python -m diff_sal_b200.build` (or __graft_entry__.build()). "
                "There is no CPU / PyTorch fallback for the denoiser." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    return _lib


def test_lib():
    """The per-kernel test entries (include/diffsal_b200_test.h); used by tests/ only, never by the product path."""
    global _test_lib
    if _test_lib is None:
        lib()
        if not os.path.exists(TEST_LIB_PATH):
            raise LibraryMissing("%s not built: run `python -m diff_sal_b200.build`" % TEST_LIB_PATH)
        _test_lib = ctypes.CDLL(TEST_LIB_PATH)
    return _test_lib


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
