"""The C-ABI library builds, loads and exports every symbol include/medmamba_b200.h declares.
No compute call is made here (no GPU in the CPU suite)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "medmamba_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from medmamba_b200 import _lib
    handle = _lib.lib()
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"


def test_host_only_entry_points():
    from medmamba_b200 import _lib
    h = _lib.lib()
    assert h.mmb_abi_version() == 1
    assert h.mmb_status_string(ctypes.c_int(0)) == b"ok"
    assert b"invalid" in h.mmb_status_string(ctypes.c_int(-1))
    assert [h.mmb_ss2d_core_dt_pad(ctypes.c_int(r)) for r in (3, 6, 12, 24, 32)] == [4, 8, 12, 24, 32]
    assert h.mmb_ss2d_core_dt_pad(ctypes.c_int(33)) == -2
    # null pointers are rejected with a status, nothing is launched
    assert h.mmb_ss2d_core_fwd(*([None] * 7), *([ctypes.c_int(1)] * 7), None) == -1


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "medmamba_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_cuda_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from medmamba_b200 import selective_scan_fn
    z = torch.zeros
    with pytest.raises(RuntimeError):
        selective_scan_fn(z(1, 4, 3), z(1, 4, 3), -torch.ones(4, 16), z(1, 4, 16, 3), z(1, 4, 16, 3))
