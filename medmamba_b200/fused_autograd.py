"""Autograd for the fused SS2D path: every forward kernel of ``ops`` paired with its hand-written backward
(``mmb_*_bwd`` in include/medmamba_b200.h).  Parameter gradients arrive as partial sums (per CTA / per batch
element / per channel tile) and are added here with ``torch.sum`` -- deterministic, no atomics."""
from __future__ import annotations

import ctypes

import torch

from . import ops
from ._lib import check, dtype_code, i64, lib, ptr, stream_ptr, timed_launch

_c_int = ctypes.c_int


def _partial_blocks() -> int:
    return lib().mmb_partial_blocks()


def _conv_partial_blocks() -> int:
    return lib().mmb_dwconv_partial_blocks()


class DwConvSiluFn(torch.autograd.Function):
    """xc = silu(dwconv3x3(x) + bias), channels-last (MedMamba.py:294-295)."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_dtype):
        xc = ops.dwconv3x3_silu(x, weight, bias, out_dtype=out_dtype)
        ctx.save_for_backward(x, weight, bias)
        return xc

    @staticmethod
    def backward(ctx, dxc):
        x, weight, bias = ctx.saved_tensors
        B, H, W, D = x.shape
        dev = x.device
        xv, px, bs = ops._token_view(x, uniform_batch=False)
        w = weight.detach().float().contiguous()
        bb = bias.detach().float().contiguous() if bias is not None else None
        g = dxc.float().contiguous()
        ds = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
        part = torch.empty((_conv_partial_blocks(), D, 10), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            with timed_launch("dwconv3x3_silu_bwd_ds", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_dwconv3x3_silu_bwd_ds(ptr(xv), ptr(w), ptr(bb), ptr(g), None, None, ptr(ds), ptr(part),
                                                     _c_int(B), _c_int(H), _c_int(W), _c_int(D), i64(px), i64(bs),
                                                     _c_int(dtype_code(xv)), stream_ptr(dev))
            check(st, "mmb_dwconv3x3_silu_bwd_ds")
            dx = torch.empty((B, H, W, D), dtype=x.dtype, device=dev)
            with timed_launch("dwconv3x3_bwd_dx", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_dwconv3x3_bwd_dx(ptr(ds), ptr(w), ptr(dx), _c_int(B), _c_int(H), _c_int(W), _c_int(D),
                                                i64(D), _c_int(dtype_code(dx)), stream_ptr(dev))
            check(st, "mmb_dwconv3x3_bwd_dx")
        tot = part.sum(0)                                   # (D, 10)
        dw = tot[:, :9].reshape(D, 1, 3, 3).to(weight.dtype)
        db = tot[:, 9].to(bias.dtype) if bias is not None else None
        return dx, dw, db, None


class _MmF32Out(torch.autograd.Function):
    """bf16 x bf16 -> fp32 GEMM (tensor cores, fp32 accumulate AND output) with its two backward GEMMs."""

    @staticmethod
    def forward(ctx, a, w):                                  # a (M, K) bf16, w (Nout, K) fp32 parameter-side
        wb = w.to(torch.bfloat16)
        ctx.save_for_backward(a, wb)
        ctx.w_dtype = w.dtype
        return torch.mm(a, wb.t(), out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        a, wb = ctx.saved_tensors
        gb = g.to(torch.bfloat16)
        da = torch.mm(gb, wb)
        dw = torch.mm(gb.t(), a, out_dtype=torch.float32).to(ctx.w_dtype)
        return da, dw


class CoreNormGateFn(torch.autograd.Function):
    """(xc, proj, z, params) -> out_norm(sum of the four directional scans) * silu(z)
    = MedMamba.py:256-286 (without x_proj) + :298-301, forward and backward in four kernels."""

    @staticmethod
    def forward(ctx, xc, proj, z, Wdt, dt_bias, A, Ds, gamma, beta, eps, d_state, dt_rank):
        Wdt_c, b_c, A_c, D_c = (Wdt.float().contiguous(), dt_bias.float().contiguous(), A.float().contiguous(),
                                Ds.float().contiguous())
        ydir, hsave = ops.ss2d_core(xc, proj, Wdt_c, b_c, A_c, D_c, d_state, dt_rank, save_states=True)
        y, merged = ops.outnorm_gate(ydir, z, gamma, beta, eps, want_merged=True, xc=xc, Ds=D_c)
        ctx.save_for_backward(xc, proj, z, Wdt_c, b_c, A_c, D_c, gamma, beta, merged, hsave)
        ctx.meta = (eps, d_state, dt_rank, Wdt.dtype, dt_bias.dtype, A.dtype, Ds.dtype)
        return y

    @staticmethod
    def backward(ctx, dout):
        xc, proj, z, Wdt, dt_bias, A, Ds, gamma, beta, merged, hsave = ctx.saved_tensors
        eps, N, R, wdt_t, b_t, a_t, d_t = ctx.meta
        B, H, W, D = xc.shape
        dev = xc.device
        f32 = dict(dtype=torch.float32, device=dev)
        rp = ops.dt_pad(R)
        zv, z_px, _ = ops._token_view(z)
        dout = dout.to(zv.dtype).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        dy = torch.empty((B, H, W, D), **f32)
        dz = torch.empty((B, H, W, D), dtype=zv.dtype, device=dev)
        gb_part = torch.empty((_partial_blocks(), 2, D), **f32)
        with torch.cuda.device(dev):
            with timed_launch("outnorm_gate_bwd", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_outnorm_gate_bwd(ptr(dout), ptr(merged), ptr(zv), ptr(g32), ptr(b32), ptr(dy), ptr(dz),
                                                ptr(gb_part), i64(B * H * W), _c_int(D), i64(z_px), i64(D), ctypes.c_float(eps),
                                                _c_int(dtype_code(zv)), stream_ptr(dev))
            check(st, "mmb_outnorm_gate_bwd")
        dudir, dproj, dA, dWdt, dD, db = core_bwd(xc, proj, dy, Wdt, dt_bias, A, Ds, hsave, N, R)
        gb = gb_part.sum(0)
        dxc = dudir.sum(3).to(xc.dtype)
        return (dxc, dproj, dz, dWdt.to(wdt_t), db.to(b_t), dA.to(a_t), dD.to(d_t), gb[0].to(gamma.dtype),
                gb[1].to(beta.dtype), None, None, None)


def core_bwd(xc, proj, dy, Wdt, dt_bias, A, Ds, hsave, d_state: int, dt_rank: int):
    """mmb_ss2d_core_bwd with its partial sums added up (deterministic: ``torch.sum`` over the partial axis).
    xc (B, H, W, D) fp32 / bf16, proj (B, H, W, 4, 32+RP) fp32, dy (B, H, W, D) fp32 = gradient of the merged sum,
    hsave as written by ``ops.ss2d_core(..., save_states=True)``  ->
    dudir (B, H, W, 4, D), dproj (B, H, W, 4, 32+RP), dA (4D, N), dWdt (4, D, R), dDs (4D,), d dt_bias (4, D), fp32."""
    B, H, W, D = xc.shape
    dev = xc.device
    N, R = d_state, dt_rank
    f32 = dict(dtype=torch.float32, device=dev)
    rp = ops.dt_pad(R)
    if proj.dtype != torch.float32 or dy.dtype != torch.float32 or xc.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("core_bwd: proj and dy must be float32, xc float32 or bfloat16")
    tiles = lib().mmb_ss2d_core_bwd_tiles(_c_int(B), _c_int(D))
    dudir = torch.empty((B, H, W, 4, D), **f32)
    dproj_p = torch.empty((tiles, B, H, W, 4, 32 + rp), **f32)
    dA_p = torch.empty((B, 4 * D, N), **f32)
    dW_p = torch.empty((B, 4 * D, rp), **f32)
    dD_p = torch.empty((B, 4 * D), **f32)
    db_p = torch.empty((B, 4 * D), **f32)
    with torch.cuda.device(dev), timed_launch("ss2d_core_bwd", f"B={B},L={H * W},D={D},R={R}"):
        st = lib().mmb_ss2d_core_bwd(ptr(xc), ptr(proj), ptr(dy), ptr(Wdt), ptr(dt_bias), ptr(A), ptr(Ds),
                                     ptr(hsave), ptr(dudir), ptr(dproj_p), ptr(dA_p), ptr(dW_p), ptr(dD_p),
                                     ptr(db_p), _c_int(B), _c_int(H), _c_int(W), _c_int(D), _c_int(N), _c_int(R), _c_int(rp),
                                     _c_int(dtype_code(xc)), stream_ptr(dev))
    check(st, "mmb_ss2d_core_bwd")
    dproj = dproj_p[0] if tiles == 1 else dproj_p.sum(0)
    return (dudir, dproj, dA_p.sum(0), dW_p.sum(0)[:, :R].reshape(4, D, R), dD_p.sum(0), db_p.sum(0).view(4, D))


class SS2DInnerFn(torch.autograd.Function):
    """Everything between in_proj and out_proj (MedMamba.py:292-301) as ONE autograd node, so that the backward's glue
    runs inside the kernels instead of as elementwise launches: the gradient of xc -- four direction slices from the
    core backward plus the x_proj GEMM's input gradient -- is summed by the dwconv backward kernel as it reads it, and
    dx / dz are written straight into the two halves of d(xz).  (As separate nodes autograd needed, per block: a
    reduction over the direction axis, an accumulate-add, a cast, two zero-filled (B, H, W, 2D) buffers, two slice
    copies and an add: ~2 GB of traffic at stage 1, batch 128.)"""

    @staticmethod
    def forward(ctx, xz, conv_w, conv_b, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, norm_w, norm_b,
                eps, d_state, dt_rank):
        B, H, W, D2 = xz.shape
        D = D2 // 2
        x, z = xz[..., :D], xz[..., D:]
        bf16 = xz.dtype == torch.bfloat16 and D % 8 == 0
        w_packed = ops.pack_x_proj(x_proj_weight.detach().float(), d_state, dt_rank)
        xc = ops.dwconv3x3_silu(x, conv_w, conv_b, out_dtype=torch.bfloat16 if bf16 else torch.float32)
        with torch.autocast("cuda", enabled=False):
            if bf16:
                wb = w_packed.to(torch.bfloat16)
                proj = torch.mm(xc.view(-1, D), wb.t(), out_dtype=torch.float32).view(B, H, W, 4, -1)
            else:
                wb = w_packed
                proj = (xc.view(-1, D) @ w_packed.t()).view(B, H, W, 4, -1)
        A = -torch.exp(A_logs.detach().float())
        Wdt_c, b_c, A_c, D_c = (dt_projs_weight.detach().float().contiguous(), dt_projs_bias.detach().float().contiguous(),
                                A.contiguous(), Ds.detach().float().contiguous())
        ydir, hsave = ops.ss2d_core(xc, proj, Wdt_c, b_c, A_c, D_c, d_state, dt_rank, save_states=True)
        y, merged = ops.outnorm_gate(ydir, z, norm_w, norm_b, eps, want_merged=True, xc=xc, Ds=D_c)
        ctx.save_for_backward(xz, conv_w, conv_b, xc, proj, wb, Wdt_c, b_c, A_c, D_c, norm_w, norm_b, merged, hsave)
        ctx.meta = (eps, d_state, dt_rank, x_proj_weight.dtype, dt_projs_weight.dtype, dt_projs_bias.dtype, A_logs.dtype,
                    Ds.dtype, x_proj_weight.shape)
        return y

    @staticmethod
    def backward(ctx, dout):
        xz, conv_w, conv_b, xc, proj, wb, Wdt, dt_bias, A, Ds, gamma, beta, merged, hsave = ctx.saved_tensors
        eps, N, R, xw_t, wdt_t, b_t, alog_t, d_t, xw_shape = ctx.meta
        B, H, W, D2 = xz.shape
        D = D2 // 2
        dev = xz.device
        f32 = dict(dtype=torch.float32, device=dev)
        rp = ops.dt_pad(R)
        x, z = xz[..., :D], xz[..., D:]
        xv, x_px, x_bs = ops._token_view(x, uniform_batch=False)
        zv, z_px, _ = ops._token_view(z)
        dout = dout.to(zv.dtype).contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        dxz = torch.empty((B, H, W, D2), dtype=xz.dtype, device=dev)          # [..., :D] = dx, [..., D:] = dz
        dy = torch.empty((B, H, W, D), **f32)
        gb_part = torch.empty((_partial_blocks(), 2, D), **f32)
        with torch.cuda.device(dev):
            with timed_launch("outnorm_gate_bwd", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_outnorm_gate_bwd(ptr(dout), ptr(merged), ptr(zv), ptr(g32), ptr(b32), ptr(dy),
                                                ctypes.c_void_p(dxz.data_ptr() + D * dxz.element_size()), ptr(gb_part),
                                                i64(B * H * W), _c_int(D), i64(z_px), i64(D2), ctypes.c_float(eps),
                                                _c_int(dtype_code(zv)), stream_ptr(dev))
            check(st, "mmb_outnorm_gate_bwd")
        dudir, dproj, dA, dWdt, dD, db = core_bwd(xc, proj, dy, Wdt, dt_bias, A, Ds, hsave, N, R)
        # x_proj backward: two GEMMs (tensor cores in the autocast layout)
        dproj2 = dproj.view(B * H * W, -1)
        with torch.autocast("cuda", enabled=False):
            if xc.dtype == torch.bfloat16:
                gb16 = dproj2.to(torch.bfloat16)
                dxe = torch.mm(gb16, wb)                                                      # (BL, D) bf16
                dwp = torch.mm(gb16.t(), xc.view(-1, D), out_dtype=torch.float32)
            else:
                dxe = dproj2 @ wb
                dwp = dproj2.t() @ xc.view(-1, D)
        if dxe.dtype != xv.dtype:
            dxe = dxe.to(xv.dtype)                                 # the kernel reads it in the dtype of x
        # un-pack d Wpacked (4 * CP, D) -> d x_proj_weight (4, R + 2N, D): rows [B_n | C_n | dt_r] per direction
        dwp = dwp.view(4, 32 + rp, D)
        dxw = torch.cat((dwp[:, 32:32 + R], dwp[:, 0:N], dwp[:, 16:16 + N]), dim=1)
        w = conv_w.detach().float().contiguous()
        bb = conv_b.detach().float().contiguous() if conv_b is not None else None
        ds = torch.empty((B, H, W, D), **f32)
        part = torch.empty((_conv_partial_blocks(), D, 10), **f32)
        with torch.cuda.device(dev):
            with timed_launch("dwconv3x3_silu_bwd_ds", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_dwconv3x3_silu_bwd_ds(ptr(xv), ptr(w), ptr(bb), None, ptr(dudir), ptr(dxe), ptr(ds), ptr(part),
                                                     _c_int(B), _c_int(H), _c_int(W), _c_int(D), i64(x_px), i64(x_bs),
                                                     _c_int(dtype_code(xv)), stream_ptr(dev))
            check(st, "mmb_dwconv3x3_silu_bwd_ds")
            with timed_launch("dwconv3x3_bwd_dx", f"B={B},L={H * W},D={D}"):
                st = lib().mmb_dwconv3x3_bwd_dx(ptr(ds), ptr(w), ptr(dxz), _c_int(B), _c_int(H), _c_int(W), _c_int(D),
                                                i64(D2), _c_int(dtype_code(dxz)), stream_ptr(dev))
            check(st, "mmb_dwconv3x3_bwd_dx")
        tot = part.sum(0)
        dcw = tot[:, :9].reshape(D, 1, 3, 3).to(conv_w.dtype)
        dcb = tot[:, 9].to(conv_b.dtype) if conv_b is not None else None
        gb = gb_part.sum(0)
        dAlog = (dA * A).to(alog_t)                                   # A = -exp(A_logs): dA/dA_logs = A
        return (dxz, dcw, dcb, dxw.to(xw_t), dWdt.to(wdt_t), db.to(b_t), dAlog, dD.to(d_t), gb[0].to(gamma.dtype),
                gb[1].to(beta.dtype), None, None, None)


def ss2d_inner_train(xz, conv_w, conv_b, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, norm_w, norm_b,
                     eps, d_state, dt_rank):
    """Differentiable twin of ops.ss2d_inner (same kernels forward, hand-written kernels backward)."""
    import os
    if os.environ.get("MMB_TRAIN_FUSED_NODE", "1") != "0" and d_state <= 16:
        return SS2DInnerFn.apply(xz, conv_w, conv_b, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, norm_w,
                                 norm_b, eps, d_state, dt_rank)
    B, H, W, D2 = xz.shape
    D = D2 // 2
    x, z = xz[..., :D], xz[..., D:]
    w_packed = ops.pack_x_proj(x_proj_weight.float(), d_state, dt_rank)
    if xz.dtype == torch.bfloat16 and D % 8 == 0:
        xc = DwConvSiluFn.apply(x, conv_w, conv_b, torch.bfloat16)
        proj = _MmF32Out.apply(xc.view(-1, D), w_packed).view(B, H, W, 4, -1)
    else:
        xc = DwConvSiluFn.apply(x, conv_w, conv_b, torch.float32)
        with torch.autocast("cuda", enabled=False):      # fp32 x_proj whatever the caller's autocast dtype is
            proj = (xc.view(-1, D) @ w_packed.t()).view(B, H, W, 4, -1)
    A = -torch.exp(A_logs.float())
    return CoreNormGateFn.apply(xc, proj, z, dt_projs_weight, dt_projs_bias, A, Ds.float(), norm_w, norm_b, eps,
                                d_state, dt_rank)


class ShuffleCatResidualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, left, ssm, inp):
        ctx.meta = (left.dtype, ssm.dtype, inp.dtype, ssm.shape)
        return ops.shuffle_cat_residual_raw(left, ssm, inp)

    @staticmethod
    def backward(ctx, g):
        ldt, sdt, idt, (B, H, W, c) = ctx.meta
        dev = g.device
        g = g.contiguous()
        bdt = sdt if sdt in (torch.float32, torch.bfloat16) and (sdt == g.dtype or g.dtype == torch.float32) else g.dtype
        dleft = torch.empty((B, H, W, c), dtype=bdt, device=dev)
        dssm = torch.empty((B, H, W, c), dtype=bdt, device=dev)
        with torch.cuda.device(dev), timed_launch("shuffle_cat_residual_bwd", f"B={B},L={H * W},c={c}"):
            st = lib().mmb_shuffle_cat_residual_bwd(ptr(g), ptr(dleft), ptr(dssm), i64(B * H * W), _c_int(c),
                                                    _c_int(dtype_code(g)), _c_int(dtype_code(dleft)), stream_ptr(dev))
        check(st, "mmb_shuffle_cat_residual_bwd")
        return dleft.to(ldt), dssm.to(sdt), g.to(idt)


class LayerNormFn(torch.autograd.Function):
    """LayerNorm over the channels of a channels-last (B, H, W, C) view with the hand-written forward and
    backward kernels (ln_1 MedMamba.py:351, the patch norms :75 and :116, in training)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        y = ops.layernorm(x, weight, bias, eps, out_dtype=out_dtype)
        ctx.save_for_backward(x, weight)
        ctx.eps = eps
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, H, W, C = x.shape
        dev = x.device
        xv, px, _ = ops._token_view(x)
        dy = dy.contiguous()
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        g32 = weight.detach().float().contiguous()
        dx = torch.empty((B, H, W, C), dtype=xv.dtype, device=dev)
        part = torch.empty((_partial_blocks(), 2, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), timed_launch("layernorm_bwd", f"B={B},L={H * W},C={C}"):
            st = lib().mmb_layernorm_bwd(ptr(xv), ptr(dy), ptr(g32), ptr(dx), ptr(part), i64(B * H * W), _c_int(C),
                                         i64(px), ctypes.c_float(ctx.eps), _c_int(dtype_code(xv)),
                                         _c_int(dtype_code(dy)), stream_ptr(dev))
        check(st, "mmb_layernorm_bwd")
        gb = part.sum(0)
        return dx, gb[0].to(ctx.param_dtypes[0]), gb[1].to(ctx.param_dtypes[1]), None, None


def train_layernorm_ok(x: torch.Tensor, ln) -> bool:
    """The LayerNormFn pair applies: CUDA, autograd needed, plain affine LayerNorm over C % 4 == 0 <= 512, no hooks."""
    return (x.is_cuda and x.dim() == 4 and isinstance(ln, torch.nn.LayerNorm) and ln.elementwise_affine
            and ln.bias is not None and len(ln.normalized_shape) == 1 and x.shape[-1] % 4 == 0 and x.shape[-1] <= 512
            and x.dtype in (torch.float32, torch.bfloat16) and ops.needs_autograd(x, ln.weight, ln.bias)
            and not ops.has_hooks(ln) and ops.fused_available())


def layernorm_train(x: torch.Tensor, ln, out_dtype=None) -> torch.Tensor:
    return LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps, out_dtype or x.dtype)
