"""The C-ABI library loads on a CPU-only box and exports every symbol include/lavila_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lavila_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lv_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from lavila_b200 import build, _lib
    build.build()
    lib = _lib.lib()
    assert lib.lv_version() == 1
    names = _declared()
    assert len(names) >= 25, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.lv_launch_count() == 0
    assert lib.lv_last_error() is not None


def test_decl_table_matches_header():
    from lavila_b200 import _decl
    declared = set(_declared())
    assert set(_decl.SIGNATURES) <= declared, set(_decl.SIGNATURES) - declared
    known = set(_decl.SIGNATURES) | {"lv_version", "lv_last_error", "lv_launch_count", "lv_gemm_bf16", "lv_gemm_bf16_2cta",
                                         "lv_gemm_skinny_bf16", "lv_gemm_skinny_splits", "lv_workspace_bytes"}   # declared in _lib.py (struct args)
    assert declared <= known, declared - known


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are rejected before any kernel launch."""
    import torch
    from lavila_b200 import ops, _lib
    a = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.LavilaB200Error):
        ops.gemm(a, a, 4, 4, 64, torch.zeros(4, 4))


def test_oracle_is_not_imported_by_the_product():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import lavila_b200, lavila_b200.engine, lavila_b200.models.models, "
            "lavila_b200.models.loss, lavila_b200.models.narrator, lavila_b200.models.gpt2_gated; "
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; assert not bad, bad" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_workspace_bytes():
    """lv_workspace_bytes (the caller owns every buffer) -- pure host arithmetic, no device needed."""
    from lavila_b200 import _lib
    lib = _lib.lib()
    s = lib.lv_gemm_skinny_splits(32, 1600, 1600)
    assert 1 <= s <= 25
    assert lib.lv_workspace_bytes(1, 32, 1600, 1600) == s * 32 * 1600 * 4
    assert lib.lv_workspace_bytes(1, 32, 1000, 1600) == -1          # N % 64 != 0: shape not supported by the skinny kernel
    assert lib.lv_workspace_bytes(2, 64, 12, 0) == 64 * 12 * 2 * 64 * 4
    assert lib.lv_workspace_bytes(3, 512, 0, 0) == 512 * 16 + 16 + 24
    assert lib.lv_workspace_bytes(4, 64, 256, 0) == 2 * (64 * 512 + 32) * 4
    assert lib.lv_workspace_bytes(99, 1, 1, 1) == -1 and lib.lv_workspace_bytes(1, 0, 64, 64) == -1


def test_error_contract_without_device():
    """Invalid arguments are refused BEFORE any CUDA call: return code < 0, message through lv_last_error(), no exception
    across the C boundary.  Runs on a box without a GPU."""
    import ctypes
    from lavila_b200 import _lib
    lib = _lib.lib()
    e = _lib.LvGemmEpilogue()
    assert lib.lv_gemm_bf16(None, 0, 0, None, 0, 0, 4, 4, 64, 1, ctypes.byref(e), None) == -1
    assert b"lv_gemm_bf16" in lib.lv_last_error()
    assert lib.lv_layernorm_fwd(None, 0, None, None, 1e-5, None, 0, None, 0, 4, 64, None) == -1
    assert b"null pointer" in lib.lv_last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.lv_layernorm_fwd(p, 0, p, p, 1e-5, p, 0, None, 0, 1, 63, None) == -1           # D % 4 != 0
    assert b"multiple of 4" in lib.lv_last_error()
    assert lib.lv_group_attn_fwd(None, 0, None, 0, None, 0, 1, 1, 1, 1, 0, None) == -1
    assert lib.lv_flash_attn_fwd(None, 0, 0, None, None, 0, 0, 64, None, 0, 1, 1, 1, 1, 0, 0.125, None) == -1
    assert lib.lv_clip_loss_fwd(None, None, None, 4, 64, None, None, None, None, None, None) == -1
    assert lib.lv_gemm_skinny_bf16(p, 64, p, 64, 4, 60, 64, p, 1, ctypes.byref(e), None) == -1   # N % 64 != 0 (and out missing)
    with pytest.raises(_lib.LavilaB200Error):
        _lib.check(-1, "probe")
