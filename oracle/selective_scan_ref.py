"""CPU restatement of ``mamba_ssm.ops.selective_scan_interface.selective_scan_ref``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The algorithm is the one whose text the
reference keeps at ``temp.py:57-139`` (third-party ``mamba_ssm==1.0.1``, README.md:19):

* ``temp.py:58-64``   casts to fp32, ``delta + delta_bias``, optional softplus
* ``temp.py:65-78``   shapes, ``x = zeros(batch, dim, dstate)``
* ``temp.py:88-98``   ``deltaA = exp(delta ⊗ A)``; ``deltaB_u = delta·B·u`` with the grouped
                      ``B G N L -> B (G H) N L`` repeat
* ``temp.py:111-125`` the L-step recurrence, ``y = <x, C_t>``, ``last_state``
* ``temp.py:135-138`` ``+ u·D``, ``· silu(z)``, cast back to the input dtype

Real ``A`` only (``MedMamba.py:28`` asserts complex away).  Written to stream over L so
that batch-64 stage shapes do not need the 10 GB ``(B, D, L, N)`` temporaries of the
einsum form; the arithmetic per element is the same sequence of fp32 operations.

Also holds the analytic O(L) backward in float64, which is the
gradient oracle: autograd through the loop above is O(L^2) in time on CPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _expand_groups(M: torch.Tensor, dim: int) -> torch.Tensor:
    """(B, G, N, L) -> (B, dim, N, L) view-by-repeat; (B, N, L) -> (B, 1, N, L) broadcast."""
    if M.dim() == 3:
        return M[:, None]
    G = M.shape[1]
    assert dim % G == 0, "dim must be a multiple of the number of B/C groups (temp.py:95)"
    return M.repeat_interleave(dim // G, dim=1)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False, compute_dtype=torch.float32):
    """u, delta: (B, D, L); A: (D, N); B, C: (B, N, L) or (B, G, N, L); D, delta_bias: (D,);
    z: (B, D, L).  Returns out (B, D, L) in ``u.dtype`` [, last_state (B, D, N) fp32].

    ``compute_dtype`` is fp32 as in the reference (``.float()`` at temp.py:59-60); tests pass
    float64 to get a high-precision truth and a differentiable fp64 path."""
    dtype_in = u.dtype
    cd = compute_dtype
    u = u.to(cd)
    delta = delta.to(cd)
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].to(cd)
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, dstate = u.shape[0], A.shape[0], A.shape[1]
    assert not A.is_complex(), "complex A is not on MedMamba's path (MedMamba.py:28)"
    assert B.dim() >= 3 and C.dim() >= 3, "MedMamba passes input-dependent B and C"
    A = A.to(cd)
    Bf = _expand_groups(B.to(cd), dim)           # (B, D|1, N, L)
    Cf = _expand_groups(C.to(cd), dim)
    x = A.new_zeros((batch, dim, dstate))
    L = u.shape[2]
    du = delta * u
    last_state = None
    ys = []
    for i in range(L):
        dA = torch.exp(delta[:, :, i, None] * A)                       # (B, D, N)
        x = dA * x + du[:, :, i, None] * Bf[:, :, :, i]
        ys.append((x * Cf[:, :, :, i]).sum(-1))
        if i == L - 1:
            last_state = x
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D.to(cd)[:, None]
    if z is not None:
        out = out * F.silu(z.to(cd))
    out = out.to(dtype=dtype_in)
    return out if not return_last_state else (out, last_state)


def selective_scan_bwd_ref(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout):
    """Analytic O(L) backward in float64 (validated against autograd
    through ``selective_scan_ref`` in tests/test_oracle.py).

    Returns dict(du, ddelta, dA, dB, dC, dD, dz, ddelta_bias, out); dB/dC have the shape of
    B/C (grouped gradients are summed over the channels of the group)."""
    f64 = torch.float64
    u = u.to(f64); draw = delta.to(f64); A = A.to(f64); dout = dout.to(f64)
    batch, dim, L = u.shape
    N = A.shape[1]
    grouped = B.dim() == 4
    G = B.shape[1] if grouped else 1
    Bf = _expand_groups(B.to(f64), dim).expand(batch, dim, N, L)
    Cf = _expand_groups(C.to(f64), dim).expand(batch, dim, N, L)
    xraw = draw + (delta_bias.to(f64)[None, :, None] if delta_bias is not None else 0.0)
    dl = F.softplus(xraw) if delta_softplus else xraw
    # forward, keeping h_{t-1}
    h = u.new_zeros(batch, dim, N)
    hs_prev = u.new_empty(batch, dim, L, N)
    y = u.new_empty(batch, dim, L)
    for t in range(L):
        hs_prev[:, :, t] = h
        a = torch.exp(dl[:, :, t, None] * A)
        h = a * h + (dl[:, :, t] * u[:, :, t])[..., None] * Bf[:, :, :, t]
        y[:, :, t] = (h * Cf[:, :, :, t]).sum(-1)
    out_pre = y + (u * D.to(f64)[:, None] if D is not None else 0.0)
    if z is not None:
        zf = z.to(f64)
        sig = torch.sigmoid(zf)
        out = out_pre * zf * sig
        dz = dout * out_pre * sig * (1 + zf * (1 - sig))
        dy = dout * zf * sig
    else:
        out, dz, dy = out_pre, None, dout
    du = dy * D.to(f64)[:, None] if D is not None else torch.zeros_like(u)
    dD = (dy * u).sum((0, 2)) if D is not None else None
    ddl = torch.zeros_like(u)
    dA = torch.zeros_like(A)
    dBf = u.new_zeros(batch, dim, N, L)
    dCf = u.new_zeros(batch, dim, N, L)
    g = u.new_zeros(batch, dim, N)           # a_{t+1} * g_{t+1}
    for t in range(L - 1, -1, -1):
        a = torch.exp(dl[:, :, t, None] * A)
        hp = hs_prev[:, :, t]
        ht = a * hp + (dl[:, :, t] * u[:, :, t])[..., None] * Bf[:, :, :, t]
        gt = dy[:, :, t, None] * Cf[:, :, :, t] + g
        dCf[:, :, :, t] = dy[:, :, t, None] * ht
        dBf[:, :, :, t] = gt * (dl[:, :, t] * u[:, :, t])[..., None]
        du[:, :, t] += dl[:, :, t] * (gt * Bf[:, :, :, t]).sum(-1)
        ddl[:, :, t] = (gt * (Bf[:, :, :, t] * u[:, :, t, None] + hp * a * A)).sum(-1)
        dA += (gt * hp * a * dl[:, :, t, None]).sum(0)
        g = a * gt
    ddraw = ddl * torch.sigmoid(xraw) if delta_softplus else ddl
    dbias = ddraw.sum((0, 2)) if delta_bias is not None else None
    if grouped:
        dB = dBf.view(batch, G, dim // G, N, L).sum(2)
        dC = dCf.view(batch, G, dim // G, N, L).sum(2)
    else:
        dB = dBf.sum(1)
        dC = dCf.sum(1)
    return dict(du=du, ddelta=ddraw, dA=dA, dB=dB, dC=dC, dD=dD, dz=dz,
                ddelta_bias=dbias, out=out)
