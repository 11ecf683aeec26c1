"""Analytic MAC model of the selective scan, mirroring the reference's own counter
(``flops_selective_scan_ref``, MedMamba.py:18-50): fvcore-style, one multiply-add = one flop,
einsum costs from ``np.einsum_path``.  Used only for reporting."""
from __future__ import annotations

import numpy as np


def _einsum_macs(shapes, equation) -> float:
    report = np.einsum_path(equation, *[np.zeros(s) for s in shapes], optimize="optimal")[1]
    for line in report.split("\n"):
        if "optimized flop" in line.lower():
            return float(np.floor(float(line.split(":")[-1]) / 2))
    raise RuntimeError("einsum_path report has no flop line")


def flops_selective_scan_ref(B=1, L=256, D=768, N=16, with_D=True, with_Z=False, with_Group=True,
                             with_complex=False):
    assert not with_complex
    total = _einsum_macs([[B, D, L], [D, N]], "bdl,dn->bdln")
    if with_Group:
        total += _einsum_macs([[B, D, L], [B, N, L], [B, D, L]], "bdl,bnl,bdl->bdln")
        per_step = B * D * N + _einsum_macs([[B, D, N], [B, D, N]], "bdn,bdn->bd")
    else:
        total += _einsum_macs([[B, D, L], [B, D, N, L], [B, D, L]], "bdl,bdnl,bdl->bdln")
        per_step = B * D * N + _einsum_macs([[B, D, N], [B, N]], "bdn,bn->bd")
    total += L * per_step
    if with_D:
        total += B * D * L
    if with_Z:
        total += B * D * L
    return total
