"""Seeded input families for the scan boundary."""
import math

import torch

STAGE_SHAPES = [(384, 3136), (768, 784), (1536, 196), (3072, 49)]   # (K*D, L), MedMamba-T at 224^2


def make_scan_inputs(family, batch, KD, L, N=16, G=4, seed=0, with_z=False, layout="NL"):
    """CPU fp32 tensors.  family: 'model' (distributions observed at random init)
    or 'stress' (unit-scale everything, non-integer A).  layout 'LN' returns B/C as views whose
    innermost stride is not 1, like the reference call site (MedMamba.py:261,267-268)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    dt = torch.exp(torch.rand(KD, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
    bias = dt + torch.log(-torch.expm1(-dt))
    if family == "model":
        u = 0.1 * rn(batch, KD, L)
        delta = 0.03 * rn(batch, KD, L)
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(KD, 1)
        D = torch.ones(KD)
        sc = 0.05
    elif family == "stress":
        u = rn(batch, KD, L)
        delta = rn(batch, KD, L)
        A = -torch.exp(rn(KD, N))
        D = rn(KD)
        sc = 1.0
    else:
        raise ValueError(family)
    if layout == "NL":
        Bm, Cm = sc * rn(batch, G, N, L), sc * rn(batch, G, N, L)
    else:
        R = 3
        xdbl = sc * rn(batch, G, L, R + 2 * N)
        Bm = xdbl[..., R:R + N].permute(0, 1, 3, 2)
        Cm = xdbl[..., R + N:].permute(0, 1, 3, 2)
    z = rn(batch, KD, L) if with_z else None
    return dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=D, z=z, delta_bias=bias)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(got, want, rtol, atol, what=""):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        worst = (err / tol).max().item()
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.numel()} elements outside rtol={rtol} atol={atol}; "
                             f"worst ratio {worst:.2f}; first at {i}: got {got[tuple(i)].item()} want {want[tuple(i)].item()}")
