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
                                         "lv_gemm_skinny_bf16", "lv_gemm_skinny_splits"}   # declared in _lib.py (struct args)
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
