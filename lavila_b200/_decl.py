"""argtypes for the non-GEMM entry points (kept next to _lib.py; grows with include/lavila_b200.h)."""
import ctypes

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


def declare(lib):
    pass
