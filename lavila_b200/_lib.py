"""ctypes binding of liblavila_b200.so -- the C ABI declared in include/lavila_b200.h.

This is the stub a reference maintainer would add.  There is NO fallback: if the shared
library is missing or a call fails the error is raised, never routed to PyTorch or the CPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblavila_b200.so")

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# epilogue flags (include/lavila_b200.h)
EPI_BIAS, EPI_QUICKGELU, EPI_DQUICKGELU, EPI_SCALE, EPI_SCALE_TANH = 1, 2, 4, 8, 16
EPI_RESID, EPI_OUT_F32, EPI_COPY_BF16, EPI_ATOMIC, EPI_ROWBIAS = 32, 64, 128, 256, 512
EPI_GELU_TANH, EPI_SQRELU = 1024, 2048


class LvGemmEpilogue(ctypes.Structure):
    _fields_ = [
        ("flags", ctypes.c_int32), ("_pad", ctypes.c_int32),
        ("out", c_void_p), ("ldo", c_int64),
        ("out2", c_void_p), ("ldo2", c_int64),
        ("bias", c_void_p),
        ("resid", c_void_p), ("ldr", c_int64),
        ("aux", c_void_p), ("ldaux", c_int64),
        ("scale_ptr", c_void_p),
    ]


class LavilaB200Error(RuntimeError):
    pass


_lib = None


def _declare(lib):
    lib.lv_version.restype = c_int
    lib.lv_last_error.restype = ctypes.c_char_p
    lib.lv_launch_count.restype = c_int64
    lib.lv_workspace_bytes.restype = c_int64
    lib.lv_workspace_bytes.argtypes = [c_int, c_int64, c_int64, c_int64]
    lib.lv_gemm_bf16.restype = c_int
    lib.lv_gemm_bf16.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int,
                                 ctypes.POINTER(LvGemmEpilogue), c_void_p]
    lib.lv_gemm_bf16_2cta.restype = c_int
    lib.lv_gemm_bf16_2cta.argtypes = lib.lv_gemm_bf16.argtypes
    lib.lv_gemm_skinny_splits.restype = c_int
    lib.lv_gemm_skinny_splits.argtypes = [c_int64, c_int64, c_int64]
    lib.lv_gemm_skinny_bf16.restype = c_int
    lib.lv_gemm_skinny_bf16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int,
                                        ctypes.POINTER(LvGemmEpilogue), c_void_p]
    from . import _decl  # remaining entry points
    _decl.declare(lib)


def lib():
    """Load (once) and return the C-ABI library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LavilaB200Error(
                "liblavila_b200.so not found at %s -- run `python -m lavila_b200.build` "
                "(there is no PyTorch/CPU fallback for the hot path)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        _declare(l)
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().lv_last_error().decode(errors="replace")
        raise LavilaB200Error("%s failed (rc=%d): %s" % (what or "lavila_b200 call", rc, msg))


def launch_count():
    return int(lib().lv_launch_count())
