"""Drop-in for ``mamba_ssm.ops.selective_scan_interface.selective_scan_fn``.

Same callable, same defaults and the same tensor contract as the function the reference imports
at ``MedMamba.py:12`` and calls at ``MedMamba.py:273-279`` (shapes documented at
``temp.py:27-36``, semantics at ``temp.py:57-139``):

    u, delta : (B, D, L)          A : (D, N)        B, C : (B, N, L) or (B, G, N, L), D % G == 0
    D, delta_bias : (D,) fp32     z : (B, D, L)     -> out (B, D, L) in u.dtype [, last_state (B, D, N)]

The arithmetic runs in the hand-written sm_100a kernels behind the C ABI
(``mmb_scan_fwd`` / ``mmb_scan_bwd`` in include/medmamba_b200.h).  CUDA tensors only: there is
no CPU path, and a missing extension raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, dtype_code, i64, lib, ptr, require_cuda, stream_ptr, timed_launch

MAX_DSTATE = 16


def _rows_contiguous(t: torch.Tensor) -> torch.Tensor:
    """(B, D, L) with unit stride along L (any batch / row stride)."""
    return t if t.stride(-1) == 1 or t.shape[-1] == 1 else t.contiguous()


def _as_grouped(M: torch.Tensor, name: str) -> torch.Tensor:
    if M.dim() == 3:
        return M.unsqueeze(1)
    if M.dim() == 4:
        return M
    raise ValueError(f"{name} must be (B, N, L) or (B, G, N, L); constant B/C are not on MedMamba's path")


def _check_shapes(u, delta, A, B, C, D, z, delta_bias):
    if u.dim() != 3:
        raise ValueError("u must be (batch, dim, seqlen)")
    batch, dim, L = u.shape
    if delta.shape != u.shape:
        raise ValueError(f"delta {tuple(delta.shape)} must match u {tuple(u.shape)}")
    if A.dim() != 2 or A.shape[0] != dim:
        raise ValueError(f"A must be (dim, dstate) = ({dim}, N), got {tuple(A.shape)}")
    if A.is_complex():
        raise ValueError("complex A is not supported (MedMamba.py:28 asserts it away)")
    N = A.shape[1]
    if N > MAX_DSTATE:
        raise ValueError(f"dstate {N} > {MAX_DSTATE}: one launch of the sm_100a kernels holds at most {MAX_DSTATE} states per "
                         f"row; selective_scan_fn splits wider state spaces into groups of {MAX_DSTATE}")
    for name, M in (("B", B), ("C", C)):
        if M.shape[0] != batch or M.shape[-2] != N or M.shape[-1] != L:
            raise ValueError(f"{name} {tuple(M.shape)} does not match (batch={batch}, ..., N={N}, L={L})")
        if dim % M.shape[1] != 0:
            raise ValueError(f"dim {dim} must be a multiple of the {name} groups {M.shape[1]}")
    if B.shape[1] != C.shape[1]:
        raise ValueError("B and C must have the same number of groups")
    for name, v in (("D", D), ("delta_bias", delta_bias)):
        if v is not None and tuple(v.shape) != (dim,):
            raise ValueError(f"{name} must be ({dim},), got {tuple(v.shape)}")
    if z is not None and z.shape != u.shape:
        raise ValueError(f"z {tuple(z.shape)} must match u {tuple(u.shape)}")
    return batch, dim, L, N


def scan_chunk_len() -> int:
    return lib().mmb_scan_chunk_len()


def scan_forward(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                 want_last_state=False, want_chunk_state=False):
    """Raw forward: returns (out, last_state | None, chunk_state | None).  No autograd."""
    dev = require_cuda(u, delta, A, B, C, D, z, delta_bias)
    B = _as_grouped(B, "B")
    C = _as_grouped(C, "C")
    batch, dim, L, N = _check_shapes(u, delta, A, B, C, D, z, delta_bias)
    if delta.dtype != u.dtype:
        delta = delta.to(u.dtype)
    if z is not None and z.dtype != u.dtype:
        z = z.to(u.dtype)
    if C.dtype != B.dtype:
        C = C.to(B.dtype)
    if B.dtype not in (torch.float32, u.dtype):
        B, C = B.float(), C.float()
    u, delta = _rows_contiguous(u), _rows_contiguous(delta)
    z = _rows_contiguous(z) if z is not None else None
    A = A.float().contiguous()
    D = D.float().contiguous() if D is not None else None
    delta_bias = delta_bias.float().contiguous() if delta_bias is not None else None
    out = torch.empty((batch, dim, L), dtype=u.dtype, device=dev)
    last = torch.empty((batch, dim, N), dtype=torch.float32, device=dev) if want_last_state else None
    chunk_state = None
    if want_chunk_state and batch > 0 and L > 0:
        T = scan_chunk_len()
        chunk_state = torch.empty((batch, dim, (L + T - 1) // T, N), dtype=torch.float32, device=dev)
    if batch == 0 or L == 0:
        if last is not None:
            last.zero_()
        return out, last, chunk_state
    zs = (z.stride(0), z.stride(1)) if z is not None else (0, 0)
    with torch.cuda.device(dev), timed_launch("scan_fwd", f"B={batch},KD={dim},L={L}"):
        st = lib().mmb_scan_fwd(
            ptr(u), ptr(delta), ptr(A), ptr(B), ptr(C), ptr(D), ptr(z), ptr(delta_bias), ptr(out),
            ptr(last), ptr(chunk_state),
            ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L), ctypes.c_int(N), ctypes.c_int(B.shape[1]),
            i64(u.stride(0)), i64(u.stride(1)), i64(delta.stride(0)), i64(delta.stride(1)),
            i64(zs[0]), i64(zs[1]), i64(out.stride(0)), i64(out.stride(1)),
            i64(B.stride(0)), i64(B.stride(1)), i64(B.stride(2)), i64(B.stride(3)),
            i64(C.stride(0)), i64(C.stride(1)), i64(C.stride(2)), i64(C.stride(3)),
            ctypes.c_int(int(bool(delta_softplus))), ctypes.c_int(dtype_code(u)), ctypes.c_int(dtype_code(B)),
            stream_ptr(dev))
    check(st, "mmb_scan_fwd")
    return out, last, chunk_state


def scan_backward(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout, chunk_state):
    """Raw backward: (du, ddelta, dA, dB, dC, dD, dz, ddelta_bias); dB/dC grouped (B, G, N, L) in B.dtype."""
    dev = require_cuda(u, delta, A, B, C, D, z, delta_bias, dout)
    Bg, Cg = _as_grouped(B, "B"), _as_grouped(C, "C")
    batch, dim, L, N = _check_shapes(u, delta, A, Bg, Cg, D, z, delta_bias)
    io = u.dtype
    delta = delta.to(io) if delta.dtype != io else delta
    z = z.to(io) if (z is not None and z.dtype != io) else z
    dout = dout.to(io) if dout.dtype != io else dout
    bc_in = Bg.dtype
    if Cg.dtype != Bg.dtype:
        Cg = Cg.to(Bg.dtype)
    if Bg.dtype not in (torch.float32, io):
        Bg, Cg = Bg.float(), Cg.float()
    u, delta, dout = _rows_contiguous(u), _rows_contiguous(delta), _rows_contiguous(dout)
    z = _rows_contiguous(z) if z is not None else None
    A32 = A.float().contiguous()
    D32 = D.float().contiguous() if D is not None else None
    b32 = delta_bias.float().contiguous() if delta_bias is not None else None
    G = Bg.shape[1]
    f32 = dict(dtype=torch.float32, device=dev)
    du = torch.empty((batch, dim, L), dtype=io, device=dev)
    ddelta = torch.empty((batch, dim, L), dtype=io, device=dev)
    dz = torch.empty((batch, dim, L), dtype=io, device=dev) if z is not None else None
    tiles = lib().mmb_scan_bwd_row_tiles(ctypes.c_int(dim), ctypes.c_int(G))
    dBp = torch.empty((tiles, batch, G, N, L), **f32)
    dCp = torch.empty((tiles, batch, G, N, L), **f32)
    dAp = torch.empty((batch, dim, N), **f32)
    dDp = torch.empty((batch, dim), **f32)
    dbp = torch.empty((batch, dim), **f32)
    if batch > 0 and L > 0:
        zs = (z.stride(0), z.stride(1)) if z is not None else (0, 0)
        with torch.cuda.device(dev), timed_launch("scan_bwd", f"B={batch},KD={dim},L={L}"):
            st = lib().mmb_scan_bwd(
                ptr(u), ptr(delta), ptr(A32), ptr(Bg), ptr(Cg), ptr(D32), ptr(z), ptr(b32), ptr(dout), ptr(chunk_state),
                ptr(du), ptr(ddelta), ptr(dz), ptr(dBp), ptr(dCp), ptr(dAp), ptr(dDp), ptr(dbp),
                ctypes.c_int(batch), ctypes.c_int(dim), ctypes.c_int(L), ctypes.c_int(N), ctypes.c_int(G),
                i64(u.stride(0)), i64(u.stride(1)), i64(delta.stride(0)), i64(delta.stride(1)),
                i64(zs[0]), i64(zs[1]), i64(dout.stride(0)), i64(dout.stride(1)),
                i64(Bg.stride(0)), i64(Bg.stride(1)), i64(Bg.stride(2)), i64(Bg.stride(3)),
                i64(Cg.stride(0)), i64(Cg.stride(1)), i64(Cg.stride(2)), i64(Cg.stride(3)),
                ctypes.c_int(int(bool(delta_softplus))), ctypes.c_int(dtype_code(u)), ctypes.c_int(dtype_code(Bg)),
                stream_ptr(dev))
        check(st, "mmb_scan_bwd")
    else:
        for t in (dBp, dCp, dAp, dDp, dbp):
            t.zero_()
    dB = (dBp[0] if tiles == 1 else dBp.sum(0)).to(bc_in)
    dC = (dCp[0] if tiles == 1 else dCp.sum(0)).to(C.dtype)
    dA = dAp.sum(0).to(A.dtype)
    dD = dDp.sum(0).to(D.dtype) if D is not None else None
    dbias = dbp.sum(0).to(delta_bias.dtype) if delta_bias is not None else None
    return du, ddelta, dA, dB, dC, dD, dz, dbias


class SelectiveScanFn(torch.autograd.Function):
    """Forward saves the inputs and the per-chunk state checkpoints; backward recomputes inside
    each chunk from its checkpoint (``mmb_scan_bwd``)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        needs_grad = any(ctx.needs_input_grad)
        out, last, chunk_state = scan_forward(
            u, delta, A, B, C, D, z, delta_bias, delta_softplus,
            want_last_state=return_last_state, want_chunk_state=needs_grad)
        ctx.delta_softplus = bool(delta_softplus)
        ctx.return_last_state = bool(return_last_state)
        ctx.b_squeezed = B.dim() == 3
        ctx.c_squeezed = C.dim() == 3
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, chunk_state)
        if return_last_state:
            ctx.mark_non_differentiable(last)
            return out, last
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        u, delta, A, B, C, D, z, delta_bias, chunk_state = ctx.saved_tensors
        grads = scan_backward(u, delta, A, B, C, D, z, delta_bias, ctx.delta_softplus, dout, chunk_state)
        du, ddelta, dA, dB, dC, dD, dz, dbias = grads
        if ctx.b_squeezed and dB is not None:
            dB = dB.squeeze(1)
        if ctx.c_squeezed and dC is not None:
            dC = dC.squeeze(1)
        return du, ddelta, dA, dB, dC, dD, dz, dbias, None, None


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """See the module docstring; signature of mamba_ssm's ``selective_scan_fn`` (MedMamba.py:12)."""
    if A.dim() == 2 and A.shape[1] > MAX_DSTATE:
        return _wide_state_scan(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
    tensors = (u, delta, A, B, C, D, z, delta_bias)
    if not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)):
        # inference (also under no_grad with leaves that require grad: ctx.needs_input_grad does not see grad mode):
        # no autograd node, no state checkpoints, the cp.async forward
        out, last, _ = scan_forward(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                    want_last_state=return_last_state, want_chunk_state=False)
        return (out, last) if return_last_state else out
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def _wide_state_scan(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state):
    """dstate > 16 (the reference's kernel takes up to 256: temp.py:27-36; VSSM(d_state=None, dims=[128, ...]) gives 22).
    The states of a row never interact -- h_n follows its own recurrence and y = sum_n C_n h_n + D u (temp.py:111-131) --
    so the state axis is cut into groups of MAX_DSTATE, every group runs through the kernels as its own scan (B / C as
    strided views of the caller's tensors, D in the first group only) and the group outputs are added in fp32; the gate
    SiLU(z) is applied to the sum.  Gradients compose through autograd: one SelectiveScanFn node per group.  A
    compatibility path (ceil(N / 16) launches), not a tuned one."""
    if B.dim() not in (3, 4) or C.dim() not in (3, 4):
        raise ValueError("B and C must be (B, N, L) or (B, G, N, L); constant B/C are not on MedMamba's path")
    N = A.shape[1]
    if N > 256:
        raise ValueError(f"dstate {N} > 256 (the limit of the operator this replaces)")
    if B.shape[-2] != N or C.shape[-2] != N:
        raise ValueError(f"B {tuple(B.shape)} / C {tuple(C.shape)} do not match dstate {N}")
    io = u.dtype
    # 16-bit callers: the group outputs would each be rounded to 16 bits before the sum, so the groups run with fp32 rows
    uf, df = u.float(), delta.float()
    acc, lasts = None, []
    for n0 in range(0, N, MAX_DSTATE):
        n1 = min(n0 + MAX_DSTATE, N)
        r = selective_scan_fn(uf, df, A[:, n0:n1], B[..., n0:n1, :], C[..., n0:n1, :], D if n0 == 0 else None, None,
                              delta_bias, delta_softplus, return_last_state)
        y = r[0] if return_last_state else r
        acc = y if acc is None else acc + y
        if return_last_state:
            lasts.append(r[1])
    if z is not None:
        acc = acc * torch.nn.functional.silu(z.float())
    out = acc.to(io)
    return (out, torch.cat(lasts, dim=-1)) if return_last_state else out
