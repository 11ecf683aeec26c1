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
    assert h.mmb_abi_version() == _lib.ABI_VERSION == 4
    h.mmb_source_digest.restype = ctypes.c_char_p
    from medmamba_b200 import build
    assert h.mmb_source_digest().decode() == build._digest()
    assert h.mmb_status_string(ctypes.c_int(0)) == b"ok"
    assert b"invalid" in h.mmb_status_string(ctypes.c_int(-1))
    assert [h.mmb_ss2d_core_dt_pad(ctypes.c_int(r)) for r in (3, 6, 12, 24, 32)] == [4, 8, 12, 24, 32]
    assert h.mmb_ss2d_core_dt_pad(ctypes.c_int(33)) == -2
    # null pointers are rejected with a status, nothing is launched
    assert h.mmb_ss2d_core_fwd(*([None] * 9), ctypes.c_int64(0), *([ctypes.c_int(1)] * 9), None) == -1


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


def test_core_plan_is_consistent_on_the_bench_shapes():
    """mmb_ss2d_core_plan (host only): tiles cover D, CTAs are whole warps, MedMamba-T at batch 256..1024 runs one lane
    per channel with 96 / 96 / 128 / 128 channels per CTA on a 148-SM part (the measured optimum, profiles/README.md)."""
    from medmamba_b200 import _lib
    h = _lib.lib()
    got = {}
    for batch in (1, 8, 64, 256, 1024):
        for (H, D) in ((56, 96), (28, 192), (14, 384), (7, 768), (128, 96)):
            S, CT, tiles = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            st = h.mmb_ss2d_core_plan(ctypes.c_int(batch), ctypes.c_int(H), ctypes.c_int(H), ctypes.c_int(D),
                                      ctypes.byref(S), ctypes.byref(CT), ctypes.byref(tiles))
            assert st == 0
            assert S.value in (1, 2, 4) and CT.value * tiles.value >= D and (CT.value * S.value) % 32 == 0
            assert CT.value <= 256 and CT.value * S.value <= 384 and (tiles.value - 1) * CT.value < D
            got[(batch, H, D)] = (S.value, CT.value, tiles.value)
    for batch in (256, 1024):
        plan = [got[(batch, H, D)][:2] for (H, D) in ((56, 96), (28, 192), (14, 384), (7, 768))]
        assert plan == [(1, 96), (1, 96), (1, 128), (1, 128)], got
    assert h.mmb_ss2d_core_plan(ctypes.c_int(0), ctypes.c_int(1), ctypes.c_int(1), ctypes.c_int(4), None, None, None) == -1
