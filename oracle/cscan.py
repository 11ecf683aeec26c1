"""ctypes binding of oracle/scan_ref.c -- TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libscan_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scan_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libscan_ref.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous().cpu()


def _prep(u, delta, A, B, C, D, z, delta_bias):
    u, delta, A, B, C, D, z, delta_bias = map(_f32, (u, delta, A, B, C, D, z, delta_bias))
    if B.dim() == 3:
        B = B[:, None].contiguous()
    if C.dim() == 3:
        C = C[:, None].contiguous()
    assert B.shape[1] == C.shape[1]
    return u, delta, A, B, C, D, z, delta_bias


def scan_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
             precision="f32"):
    """Returns (out, last_state) as torch CPU tensors (float32 or float64)."""
    u, delta, A, B, C, D, z, delta_bias = _prep(u, delta, A, B, C, D, z, delta_bias)
    batch, dim, L = u.shape
    N, G = A.shape[1], B.shape[1]
    dt = torch.float32 if precision == "f32" else torch.float64
    out = torch.empty(batch, dim, L, dtype=dt)
    last = torch.empty(batch, dim, N, dtype=dt)
    fn = getattr(_load(), "scan_fwd_ref_" + precision)
    rc = fn(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
            ctypes.c_int(int(delta_softplus)), _p(out), _p(last),
            *(ctypes.c_long(v) for v in (batch, dim, L, N, G)))
    if rc != 0:
        raise RuntimeError(f"scan_fwd_ref_{precision} failed: {rc}")
    return out, last


def scan_bwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout):
    """Analytic float64 gradients; dict like oracle.selective_scan_ref.selective_scan_bwd_ref."""
    b3 = B.dim() == 3
    u, delta, A, B, C, D, z, delta_bias = _prep(u, delta, A, B, C, D, z, delta_bias)
    dout = _f32(dout)
    batch, dim, L = u.shape
    N, G = A.shape[1], B.shape[1]
    f64 = torch.float64
    du = torch.empty(batch, dim, L, dtype=f64)
    ddelta = torch.empty_like(du)
    dA = torch.empty(dim, N, dtype=f64)
    dB = torch.empty(batch, G, N, L, dtype=f64)
    dC = torch.empty_like(dB)
    dD = torch.empty(dim, dtype=f64) if D is not None else None
    dz = torch.empty_like(du) if z is not None else None
    dbias = torch.empty(dim, dtype=f64) if delta_bias is not None else None
    rc = _load().scan_bwd_ref_f64(
        _p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
        ctypes.c_int(int(delta_softplus)), _p(dout), _p(du), _p(ddelta), _p(dA), _p(dB), _p(dC),
        _p(dD), _p(dz), _p(dbias), *(ctypes.c_long(v) for v in (batch, dim, L, N, G)))
    if rc != 0:
        raise RuntimeError(f"scan_bwd_ref_f64 failed: {rc}")
    if b3:
        dB, dC = dB[:, 0], dC[:, 0]
    return dict(du=du, ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, dz=dz, ddelta_bias=dbias)
