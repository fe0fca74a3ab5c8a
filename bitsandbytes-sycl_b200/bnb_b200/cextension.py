"""Loader of the native library -- the drop-in boundary.

Mirrors python_src_quants/cextension.py:67-110 of the reference: the package talks to ONE shared library
through ctypes (`lib`), a library is "GPU-capable" iff it exports `get_context`, and pointer-returning
symbols get `restype = c_void_p`.  Differences: the library is libbitsandbytes_b200.so (hand-written
CUDA for sm_100a, built in-tree by ../build.py) and there is NO CPU / other-backend fallback: if the
library cannot be loaded the import fails loudly.
"""
import ctypes as ct
import logging
import os

logger = logging.getLogger(__name__)

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libbitsandbytes_b200.so"
LIB_PATH = os.path.join(PACKAGE_DIR, LIB_NAME)


class BNBNativeLibrary:
    """Same role as the reference's BNBNativeLibrary / CudaBNBNativeLibrary pair."""

    compiled_with_cuda = True

    def __init__(self, lib: ct.CDLL):
        self._lib = lib
        lib.get_context.restype = ct.c_void_p
        lib.cbnb_get_stream.restype = ct.c_void_p
        lib.cbnb_last_error_string.restype = ct.c_char_p
        lib.cbnb_version.restype = ct.c_char_p
        lib.cbnb_selftest_quant_lut.restype = ct.c_longlong

    def __getattr__(self, item):
        return getattr(self._lib, item)


def _build_if_possible() -> None:
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bnb_b200_build", os.path.join(os.path.dirname(PACKAGE_DIR), "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def get_native_library() -> BNBNativeLibrary:
    if not os.path.exists(LIB_PATH):
        logger.warning("%s not found, building it with nvcc for sm_100a", LIB_NAME)
        _build_if_possible()
    dll = ct.cdll.LoadLibrary(LIB_PATH)  # raises OSError when missing / unloadable: no fallback
    if not hasattr(dll, "get_context"):
        raise RuntimeError(f"{LIB_PATH} is not the bnb_b200 CUDA library (no get_context symbol)")
    return BNBNativeLibrary(dll)


lib = get_native_library()
COMPILED_WITH_CUDA = True
