"""Host-side mirror of the reference's module API around the SS2D hot path.

Same class names, constructor arguments, parameter names / shapes and ``state_dict`` keys as
``MedMamba.py`` (``SS2D`` :123-191, ``SS_Conv_SSM`` :322-347, ``VSSLayer`` :359-410, ``VSSM``
:423-473, ``PatchEmbed2D`` :54-76, ``PatchMerging2D`` :79-119), so reference checkpoints load with
``load_state_dict`` and consumers that walk the module tree (``test.py:101``) keep working.
Parameter initialisation draws from the RNG in the same order as the reference, so
``torch.manual_seed(s); VSSM(...)`` yields the same random-init weights (checked in
tests/test_oracle.py against the unmodified reference).

What differs is what runs between ``in_proj`` and ``out_proj`` and after the concat: on CUDA the
hand-written sm_100a kernels behind the C ABI (``medmamba_b200.ops``).  The ``in_proj`` /
``x_proj`` / ``out_proj`` linears and the CNN branch stay ordinary torch ops (cuBLAS / cuDNN).
"""
from __future__ import annotations

import math
import os
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from . import ops
from .layers import DropPath, trunc_normal_
from .selective_scan_interface import selective_scan_fn


def _train_ln_ok(x, ln) -> bool:
    from .fused_autograd import train_layernorm_ok
    return train_layernorm_ok(x, ln)


def _train_ln(x, ln, out_dtype):
    from .fused_autograd import layernorm_train
    return layernorm_train(x, ln, out_dtype)


class PatchEmbed2D(nn.Module):
    """patch_size x patch_size strided conv, NCHW -> NHWC, optional norm (MedMamba.py:54-76)."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None, **kwargs):
        super().__init__()
        ps = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=ps, stride=ps)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        if self.norm is not None and getattr(self, "fused", True) and ops.fast_patch_embed_ok(x, self.proj, self.norm):
            return ops.patch_embed_ln(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                      self.norm.eps)
        x = self.proj(x).permute(0, 2, 3, 1)
        if self.norm is None:
            return x
        if ops.fast_layernorm_ok(x, self.norm):
            return ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype=torch.float32)
        if getattr(self, "fused", True) and _train_ln_ok(x, self.norm):
            return _train_ln(x, self.norm, torch.float32)
        return self.norm(x)


class PatchMerging2D(nn.Module):
    """2x2 neighbourhood -> channels, LayerNorm(4c), Linear(4c -> 2c) (MedMamba.py:79-119).
    Odd H / W are truncated like the reference does (MedMamba.py:96-111)."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        B, H, W, C = x.shape
        if getattr(self, "fused", True) and ops.fast_patch_merge_ok(x, self.norm):
            return self.reduction(ops.patch_merge_ln(x, self.norm.weight, self.norm.bias, self.norm.eps,
                                                     out_dtype=ops.autocast_dtype(x.dtype)))
        h2, w2 = H // 2, W // 2
        quads = [x[:, i::2, j::2, :][:, :h2, :w2, :] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))]
        x = torch.cat(quads, dim=-1).view(B, h2, w2, 4 * C)
        if ops.fast_layernorm_ok(x, self.norm):
            return self.reduction(ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps,
                                                out_dtype=ops.autocast_dtype(x.dtype)))
        if getattr(self, "fused", True) and _train_ln_ok(x, self.norm):
            return self.reduction(_train_ln(x, self.norm, ops.autocast_dtype(x.dtype)))
        return self.reduction(self.norm(x))


class SS2D(nn.Module):
    """2-D selective-scan block (MedMamba.py:123-305).

    ``fused=True`` (default) runs the B200 path: dwconv3x3+SiLU kernel (channels-last) -> one
    x_proj matmul on un-permuted tokens -> ``ss2d_core`` kernel (4-direction scan with dt_proj,
    softplus and the cross-merge folded in) -> out_norm * SiLU(z) kernel.  ``fused=False`` follows
    the reference's op sequence with ``selective_scan_fn`` at the same boundary
    (``forward_corev0``, the attribute the reference rebinds at MedMamba.py:187).
    """

    def __init__(self, d_model, d_state=16, d_conv=3, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, dropout=0.0, conv_bias=True, bias=False,
                 device=None, dtype=None, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model = d_model
        self.d_state = d_state
        self.d_conv = d_conv
        self.expand = expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        K, D, N, R = 4, self.d_inner, d_state, self.dt_rank

        self.in_proj = nn.Linear(d_model, 2 * D, bias=bias, **fk)
        self.conv2d = nn.Conv2d(D, D, kernel_size=d_conv, padding=(d_conv - 1) // 2, groups=D, bias=conv_bias, **fk)
        self.act = nn.SiLU()

        # K independent x_proj / dt_proj linears, stored stacked (MedMamba.py:164-181)
        xw = [nn.Linear(D, R + 2 * N, bias=False, **fk).weight for _ in range(K)]
        self.x_proj_weight = nn.Parameter(torch.stack(xw, dim=0))                      # (K, R+2N, D)
        dts = [self.dt_init(R, D, dt_scale, dt_init, dt_min, dt_max, dt_init_floor, **fk) for _ in range(K)]
        self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dts], dim=0))   # (K, D, R)
        self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dts], dim=0))       # (K, D)
        self.A_logs = self.A_log_init(N, D, copies=K, merge=True)                      # (K*D, N)
        self.Ds = self.D_init(D, copies=K, merge=True)                                 # (K*D,)

        self.forward_core = self.forward_corev0
        self.fused = True

        self.out_norm = nn.LayerNorm(D)
        self.out_proj = nn.Linear(D, d_model, bias=bias, **fk)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None

    # -- parameter initialisers (MedMamba.py:193-247) ---------------------------------------------
    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4,
                **factory_kwargs):
        proj = nn.Linear(dt_rank, d_inner, bias=True, **factory_kwargs)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(proj.weight, -std, std)
        else:
            raise NotImplementedError(dt_init)
        # bias = softplus^-1(dt), dt log-uniform in [dt_min, dt_max]
        span = math.log(dt_max) - math.log(dt_min)
        dt = torch.exp(torch.rand(d_inner, **factory_kwargs) * span + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        proj.bias._no_reinit = True
        return proj

    @staticmethod
    def A_log_init(d_state, d_inner, copies=1, device=None, merge=True):
        A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1)
        if copies > 1:
            A_log = A_log.unsqueeze(0).repeat(copies, 1, 1)
            if merge:
                A_log = A_log.flatten(0, 1)
        A_log = nn.Parameter(A_log.contiguous())
        A_log._no_weight_decay = True
        return A_log

    @staticmethod
    def D_init(d_inner, copies=1, device=None, merge=True):
        D = torch.ones(d_inner, device=device)
        if copies > 1:
            D = D.unsqueeze(0).repeat(copies, 1)
            if merge:
                D = D.flatten(0, 1)
        D = nn.Parameter(D.contiguous())
        D._no_weight_decay = True
        return D

    # -- reference-order path: materialised cross-scan + selective_scan_fn (MedMamba.py:249-286) --
    def forward_corev0(self, x: torch.Tensor):
        self.selective_scan = selective_scan_fn
        B, C, H, W = x.shape
        L, K = H * W, 4
        R, N = self.dt_rank, self.d_state
        xs = ops.cross_scan(x)                                                        # (B, 4, C, L)
        x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, self.x_proj_weight)
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
        dts = torch.einsum("bkrl,kdr->bkdl", dts, self.dt_projs_weight)
        out_y = self.selective_scan(
            xs.float().view(B, -1, L), dts.contiguous().float().view(B, -1, L),
            -torch.exp(self.A_logs.float()).view(-1, N), Bs.float(), Cs.float(), self.Ds.float().view(-1),
            z=None, delta_bias=self.dt_projs_bias.float().view(-1), delta_softplus=True,
            return_last_state=False).view(B, K, -1, L)
        assert out_y.dtype == torch.float
        return ops.cross_merge(out_y, H, W)

    def _forward_reference_order(self, x: torch.Tensor):
        B, H, W, _ = x.shape
        xz = self.in_proj(x)
        x, z = xz.chunk(2, dim=-1)
        x = self.act(self.conv2d(x.permute(0, 3, 1, 2).contiguous()))
        y1, y2, y3, y4 = self.forward_core(x)
        assert y1.dtype == torch.float32
        y = (y1 + y2 + y3 + y4).transpose(1, 2).contiguous().view(B, H, W, -1)
        y = self.out_norm(y) * F.silu(z)
        return self.out_proj(y)

    # -- fused B200 path -----------------------------------------------------------------------------
    def _forward_fused(self, x: torch.Tensor):
        xz = self.in_proj(x)                                                          # (B, H, W, 2D)
        if xz.dtype == torch.float16:
            xz = xz.float()
        inner = ops.ss2d_inner
        if ops.needs_autograd(xz, self.conv2d.weight, self.x_proj_weight, self.dt_projs_weight, self.A_logs, self.Ds,
                              self.out_norm.weight):
            from .fused_autograd import ss2d_inner_train as inner      # same kernels + their backward kernels
        y = inner(xz, self.conv2d.weight, self.conv2d.bias, self.x_proj_weight, self.dt_projs_weight,
                           self.dt_projs_bias, self.A_logs, self.Ds, self.out_norm.weight, self.out_norm.bias,
                           self.out_norm.eps, self.d_state, self.dt_rank)
        return self.out_proj(y.to(xz.dtype) if y.dtype != xz.dtype else y)

    def forward(self, x: torch.Tensor, **kwargs):
        use_fused = (self.fused and x.is_cuda and self.d_conv == 3
                     and ops.fused_supported(self.d_state, self.dt_rank, self.d_inner) and ops.fused_available()
                     and not ops.has_hooks(self.conv2d, self.act, self.out_norm))
        out = self._forward_fused(x) if use_fused else self._forward_reference_order(x)
        return out if self.dropout is None else self.dropout(out)


def channel_shuffle(x: torch.Tensor, groups: int) -> torch.Tensor:
    """(B, H, W, C): out[..., j*groups + g] = x[..., g*(C/groups) + j] (MedMamba.py:308-320)."""
    B, H, W, C = x.shape
    return x.view(B, H, W, groups, C // groups).transpose(3, 4).reshape(B, H, W, C)


class SS_Conv_SSM(nn.Module):
    """Half the channels through LN -> SS2D, half through a small CNN, interleaved back together
    with a residual (MedMamba.py:322-357)."""

    def __init__(self, hidden_dim: int = 0, drop_path: float = 0,
                 norm_layer: Callable[..., nn.Module] = None, attn_drop_rate: float = 0, d_state: int = 16,
                 **kwargs):
        super().__init__()
        if norm_layer is None:
            norm_layer = lambda c: nn.LayerNorm(c, eps=1e-6)
        c = hidden_dim // 2
        self.ln_1 = norm_layer(c)
        self.self_attention = SS2D(d_model=c, dropout=attn_drop_rate, d_state=d_state, **kwargs)
        self.drop_path = DropPath(drop_path)
        self.conv33conv33conv11 = nn.Sequential(
            nn.BatchNorm2d(c),
            nn.Conv2d(c, c, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(c),
            nn.ReLU(),
            nn.Conv2d(c, c, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(c),
            nn.ReLU(),
            nn.Conv2d(c, c, kernel_size=1, stride=1),
            nn.ReLU(),
        )

    # -- CNN branch, inference fast path -----------------------------------------------------------------
    def _folded_cnn(self, dtype):
        """Eval-mode BatchNorms folded away (cached): the leading BN becomes a per-channel affine applied while
        gathering the left half (mmb_affine_cast_fwd); the two BNs that follow convolutions fold into those
        convolutions' weights and biases exactly.  Invalidated when any parameter / buffer changes."""
        seq = self.conv33conv33conv11
        # _version follows in-place updates (optimizer steps, load_state_dict); data_ptr / device follow rebinding
        # through `.data` (EMA weight swaps, net.to(other_device) after a forward)
        key = (dtype,) + tuple((t._version, t.data_ptr(), t.device) for m in seq
                               for t in list(m.parameters()) + list(m.buffers()))
        cached = getattr(self, "_cnn_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        with torch.no_grad():
            def bn_affine(bn):
                sc = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
                return sc, bn.bias.float() - bn.running_mean.float() * sc
            sc0, sh0 = bn_affine(seq[0])
            convs = []
            for conv, bn in ((seq[1], seq[2]), (seq[4], seq[5]), (seq[7], None)):
                w, b = conv.weight.float(), conv.bias.float()
                if bn is not None:
                    sc, sh = bn_affine(bn)
                    w, b = w * sc.view(-1, 1, 1, 1), b * sc + sh
                convs.append((w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype).contiguous(),
                              conv.padding))
            folded = (sc0.contiguous(), sh0.contiguous(), convs)
        self._cnn_cache = (key, folded)
        return folded

    def _cnn_branch_fast(self, left: torch.Tensor) -> torch.Tensor:
        dtype = ops.autocast_dtype(left.dtype)
        sc0, sh0, convs = self._folded_cnn(dtype)
        x = ops.affine_cast(left, sc0, sh0, dtype).permute(0, 3, 1, 2)        # NCHW shape, channels-last memory
        with torch.autocast("cuda", enabled=False):
            for w, b, pad in convs:
                x = torch.cudnn_convolution_relu(x, w, b, (1, 1), pad, (1, 1), 1)
        return x.permute(0, 2, 3, 1)

    @staticmethod
    def _dense_nchw_view(left: torch.Tensor) -> torch.Tensor:
        """CNN branch input in NCHW *shape* with dense channels-last memory.  `left` is the first half of the channels
        of the residual stream -- a strided view; permuted as it is, its strides are not dense, so BatchNorm takes ATen's
        generic kernels and the autocast cast falls back to NCHW memory (layout-conversion kernels before every
        convolution).  One small copy keeps the whole branch on the channels-last kernels."""
        if left.is_cuda and os.environ.get("MMB_CNN_DENSE", "1") != "0":
            left = left.contiguous()
        return left.permute(0, 3, 1, 2)

    def _cnn_fast_ok(self, left: torch.Tensor) -> bool:
        return (getattr(self, "fast_cnn", True) and left.is_cuda and not self.training and left.shape[-1] % 4 == 0
                and left.dtype in (torch.float32, torch.bfloat16) and ops.fused_available()
                and not ops.needs_autograd(left, *self.conv33conv33conv11.parameters())
                and not ops.has_hooks(self.conv33conv33conv11))

    def forward(self, input: torch.Tensor):
        left, right = input.chunk(2, dim=-1)
        if ops.fast_layernorm_ok(right, self.ln_1):
            normed = ops.layernorm(right, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps,
                                   out_dtype=ops.autocast_dtype(right.dtype))
        elif self.self_attention.fused and _train_ln_ok(right, self.ln_1):
            normed = _train_ln(right, self.ln_1, ops.autocast_dtype(right.dtype))
        else:
            normed = self.ln_1(right)
        if self._cnn_fast_ok(left) and ops.branch_overlap_enabled():
            # The two branches of a block are independent (MedMamba.py:352-355): the CNN branch (cuDNN, tensor cores
            # + HBM) runs on a side stream next to the SS2D branch (MUFU-bound scan), filling the SMs the partial
            # last round of the scan leaves idle.
            main = torch.cuda.current_stream(input.device)
            side = ops.side_stream(input.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                left = self._cnn_branch_fast(left)
            ssm = self.drop_path(self.self_attention(normed))
            main.wait_stream(side)
            # `left` lives in the side stream's allocator pool and is read by the shuffle kernel on the main stream.
            # No record_stream: the side stream is used by this method only and every use starts with
            # side.wait_stream(main), so the pool cannot hand the block out again before that kernel has run
            # (record_stream's deferred frees made the allocator grow for several steps: cudaMalloc in the loop).
            return ops.shuffle_cat_residual(left, ssm, input)
        if (self.training and input.is_cuda and torch.is_grad_enabled() and ops.train_branch_overlap_enabled()
                and not torch.cuda.is_current_stream_capturing()):
            # Training: the same fork / join around the CNN branch (cuDNN convolutions + BatchNorm).  Autograd runs every
            # node's backward on the stream its forward ran on, so the branch's dgrad / wgrad / BatchNorm backward
            # also leave the main stream and overlap the SS2D branch's backward (latency-bound, XU 29 % busy).
            main = torch.cuda.current_stream(input.device)
            side = ops.side_stream(input.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                left = self.conv33conv33conv11(self._dense_nchw_view(left)).permute(0, 2, 3, 1)
            ssm = self.drop_path(self.self_attention(normed))
            main.wait_stream(side)
            if ops.shuffle_supported(ssm.shape[-1]) and ops.fused_available():
                return ops.shuffle_cat_residual(left, ssm, input)
            return channel_shuffle(torch.cat((left, ssm), dim=-1), groups=2) + input
        ssm = self.drop_path(self.self_attention(normed))
        if self._cnn_fast_ok(left):
            left = self._cnn_branch_fast(left)
        else:
            left = self.conv33conv33conv11(self._dense_nchw_view(left)).permute(0, 2, 3, 1)
        if input.is_cuda and ops.shuffle_supported(ssm.shape[-1]) and ops.fused_available():
            return ops.shuffle_cat_residual(left, ssm, input)
        return channel_shuffle(torch.cat((left, ssm), dim=-1), groups=2) + input


class VSSLayer(nn.Module):
    """One stage: ``depth`` blocks and an optional downsample (MedMamba.py:359-422)."""

    def __init__(self, dim, depth, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None,
                 use_checkpoint=False, d_state=16, **kwargs):
        super().__init__()
        self.dim = dim
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SS_Conv_SSM(hidden_dim=dim, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                        norm_layer=norm_layer, attn_drop_rate=attn_drop, d_state=d_state)
            for i in range(depth)])
        # The reference re-initialises a *clone* of every out_proj.weight here (MedMamba.py:398-404):
        # no parameter changes, but the RNG advances.  Reproduced so seeds give identical weights.
        for blk in self.blocks:
            nn.init.kaiming_uniform_(blk.self_attention.out_proj.weight.detach().clone(), a=math.sqrt(5))
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x, use_reentrant=False) if self.use_checkpoint else blk(x)
        return x if self.downsample is None else self.downsample(x)


class VSSM(nn.Module):
    """MedMamba backbone + classifier head (MedMamba.py:423-515).  MedMamba-T is
    ``VSSM(depths=[2, 2, 4, 2], dims=[96, 192, 384, 768], num_classes=...)`` (train.py:179)."""

    def __init__(self, patch_size=4, in_chans=3, num_classes=1000, depths=[2, 2, 4, 2],
                 dims=[96, 192, 384, 768], d_state=16, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, patch_norm=True, use_checkpoint=False, **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.num_layers = len(depths)
        if isinstance(dims, int):
            dims = [int(dims * 2 ** i) for i in range(self.num_layers)]
        self.embed_dim = dims[0]
        self.num_features = dims[-1]
        self.dims = dims

        self.patch_embed = PatchEmbed2D(patch_size=patch_size, in_chans=in_chans, embed_dim=self.embed_dim,
                                        norm_layer=norm_layer if patch_norm else None)
        self.ape = False
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            lo, hi = sum(depths[:i]), sum(depths[:i + 1])
            self.layers.append(VSSLayer(
                dim=dims[i], depth=depths[i], d_state=math.ceil(dims[0] / 6) if d_state is None else d_state,
                drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[lo:hi], norm_layer=norm_layer,
                downsample=PatchMerging2D if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()

        self.apply(self._init_weights)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _init_weights(self, m: nn.Module):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"relative_position_bias_table"}

    def forward_backbone(self, x):
        x = self.pos_drop(self.patch_embed(x))
        for layer in self.layers:
            x = layer(x)
        return x

    def forward(self, x):
        x = self.forward_backbone(x)                       # (B, H, W, C)
        x = self.avgpool(x.permute(0, 3, 1, 2))
        return self.head(torch.flatten(x, start_dim=1))


def medmamba_t(num_classes=6, **kw):
    return VSSM(depths=[2, 2, 4, 2], dims=[96, 192, 384, 768], num_classes=num_classes, **kw)


def medmamba_s(num_classes=6, **kw):
    return VSSM(depths=[2, 2, 8, 2], dims=[96, 192, 384, 768], num_classes=num_classes, **kw)


def medmamba_b(num_classes=6, **kw):
    return VSSM(depths=[2, 2, 12, 2], dims=[128, 256, 512, 1024], num_classes=num_classes, **kw)
