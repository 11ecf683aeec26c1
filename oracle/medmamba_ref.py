"""Functional CPU restatement of the reference model around the SS2D hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function works on a plain ``state_dict``
whose keys are the reference's (``MedMamba.py`` module tree), so it can be fed the weights of
either the unmodified reference ``VSSM`` (container) or of ``medmamba_b200.VSSM`` (GPU box) and
is the CPU baseline ``bench.py`` times.  It restates, with torch CPU ops:

* ``cross_scan`` / ``cross_merge``      MedMamba.py:256-257, 282-286 (and, independently, the
                                        closed-form index maps of SURVEY.md Appendix A)
* ``ss2d_core``                         MedMamba.py:249-286  (forward_corev0)
* ``ss2d_forward``                      MedMamba.py:288-305
* ``channel_shuffle``                   MedMamba.py:308-320
* ``block_forward``                     MedMamba.py:349-357  (SS_Conv_SSM, eval-mode BatchNorm)
* ``patch_embed`` / ``patch_merge``     MedMamba.py:72-76, 93-119
* ``vssm_forward``                      MedMamba.py:499-515

Checked against the unmodified reference module in tests/test_oracle.py (container) and through
the golden fixtures under tests/golden/ (everywhere).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .selective_scan_ref import selective_scan_ref


# ----------------------------------------------------------------------------- index maps
def cross_scan_index(H: int, W: int) -> np.ndarray:
    """src[k, l]: flat position p = h*W + w read by direction k at sequence index l
    (the closed form of MedMamba.py:256-257)."""
    L = H * W
    l = np.arange(L)
    src = np.empty((4, L), dtype=np.int64)
    src[0] = l
    src[1] = (l % H) * W + (l // H)
    src[2] = L - 1 - l
    src[3] = src[1][::-1]
    return src


def cross_scan(x: torch.Tensor) -> torch.Tensor:
    """(B, D, H, W) -> (B, 4, D, L) exactly as MedMamba.py:256-257 does it."""
    B, D, H, W = x.shape
    L = H * W
    hw = x.reshape(B, D, L)
    wh = x.transpose(2, 3).contiguous().reshape(B, D, L)
    both = torch.stack([hw, wh], dim=1)
    return torch.cat([both, both.flip(-1)], dim=1)


def cross_merge(out_y: torch.Tensor, H: int, W: int):
    """(B, 4, D, L) -> the four (B, D, L) tensors of MedMamba.py:282-286, in return order."""
    B, K, D, L = out_y.shape
    inv = out_y[:, 2:4].flip(-1)
    wh = out_y[:, 1].reshape(B, D, W, H).transpose(2, 3).contiguous().reshape(B, D, L)
    invwh = inv[:, 1].reshape(B, D, W, H).transpose(2, 3).contiguous().reshape(B, D, L)
    return out_y[:, 0], inv[:, 0], wh, invwh


# ----------------------------------------------------------------------------- SS2D
def ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, scan_fn=None):
    """x: (B, D, H, W) after conv+SiLU.  Returns (y1, y2, y3, y4), each (B, D, L) fp32."""
    scan_fn = scan_fn or selective_scan_ref
    B, D, H, W = x.shape
    L, K = H * W, 4
    N = A_logs.shape[1]
    R = dt_projs_weight.shape[2]
    xs = cross_scan(x)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, x_proj_weight)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts, dt_projs_weight)
    out_y = scan_fn(
        xs.float().reshape(B, K * D, L), dts.contiguous().float().reshape(B, K * D, L),
        -torch.exp(A_logs.float()).reshape(K * D, N), Bs.float(), Cs.float(),
        Ds.float().reshape(-1), z=None, delta_bias=dt_projs_bias.float().reshape(-1),
        delta_softplus=True, return_last_state=False).reshape(B, K, D, L)
    return cross_merge(out_y, H, W)


def ss2d_forward(sd, p, x, scan_fn=None):
    """x: (B, H, W, d_model) -> (B, H, W, d_model); ``p`` is the key prefix of the SS2D module."""
    B, H, W, _ = x.shape
    xz = F.linear(x, sd[p + "in_proj.weight"], sd.get(p + "in_proj.bias"))
    xi, z = xz.chunk(2, dim=-1)
    D = xi.shape[-1]
    xi = xi.permute(0, 3, 1, 2).contiguous()
    xi = F.silu(F.conv2d(xi, sd[p + "conv2d.weight"], sd.get(p + "conv2d.bias"), padding=1, groups=D))
    y1, y2, y3, y4 = ss2d_core(xi, sd[p + "x_proj_weight"], sd[p + "dt_projs_weight"],
                               sd[p + "dt_projs_bias"], sd[p + "A_logs"], sd[p + "Ds"], scan_fn)
    y = y1 + y2 + y3 + y4
    y = y.transpose(1, 2).contiguous().view(B, H, W, -1)
    y = F.layer_norm(y, (D,), sd[p + "out_norm.weight"], sd[p + "out_norm.bias"], 1e-5)
    y = y * F.silu(z)
    return F.linear(y, sd[p + "out_proj.weight"], sd.get(p + "out_proj.bias"))


def channel_shuffle(x: torch.Tensor, groups: int) -> torch.Tensor:
    B, H, W, C = x.shape
    return x.view(B, H, W, groups, C // groups).transpose(3, 4).reshape(B, H, W, C)


def _bn_eval(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], training=False, eps=1e-5)


def cnn_branch(sd, p, x):
    """conv33conv33conv11 (MedMamba.py:337-347), eval mode; x: (B, c, H, W)."""
    x = _bn_eval(x, sd, p + "0.")
    x = F.conv2d(x, sd[p + "1.weight"], sd[p + "1.bias"], padding=1)
    x = F.relu(_bn_eval(x, sd, p + "2."))
    x = F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"], padding=1)
    x = F.relu(_bn_eval(x, sd, p + "5."))
    return F.relu(F.conv2d(x, sd[p + "7.weight"], sd[p + "7.bias"]))


def block_forward(sd, p, inp, scan_fn=None):
    """SS_Conv_SSM.forward in eval mode (DropPath is the identity)."""
    left, right = inp.chunk(2, dim=-1)
    c = right.shape[-1]
    r = F.layer_norm(right, (c,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
    x = ss2d_forward(sd, p + "self_attention.", r, scan_fn)
    left = cnn_branch(sd, p + "conv33conv33conv11.", left.permute(0, 3, 1, 2).contiguous())
    left = left.permute(0, 2, 3, 1).contiguous()
    return channel_shuffle(torch.cat((left, x), dim=-1), 2) + inp


def patch_embed(sd, x):
    w = sd["patch_embed.proj.weight"]
    x = F.conv2d(x, w, sd["patch_embed.proj.bias"], stride=w.shape[-1]).permute(0, 2, 3, 1)
    if "patch_embed.norm.weight" in sd:
        x = F.layer_norm(x, (x.shape[-1],), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    return x


def patch_merge(sd, p, x):
    B, H, W, C = x.shape
    h2, w2 = H // 2, W // 2
    parts = [x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]]
    parts = [t[:, :h2, :w2] for t in parts]
    x = torch.cat(parts, -1).reshape(B, h2, w2, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def vssm_forward(sd, x, depths=(2, 2, 4, 2), scan_fn=None):
    """Eval-mode VSSM.forward: x (B, 3, Hi, Wi) -> logits (B, num_classes)."""
    x = patch_embed(sd, x)
    for i, depth in enumerate(depths):
        for j in range(depth):
            x = block_forward(sd, f"layers.{i}.blocks.{j}.", x, scan_fn)
        if i < len(depths) - 1:
            x = patch_merge(sd, f"layers.{i}.downsample.", x)
    x = x.permute(0, 3, 1, 2).mean((2, 3))
    return F.linear(x, sd["head.weight"], sd["head.bias"])
