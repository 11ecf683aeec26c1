"""Generates tests/golden/*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run from the repo root:
    python -m oracle.make_golden
Everything stored is an OUTPUT OF THE REFERENCE ITSELF: /root/reference/MedMamba.py imported
unmodified, with selective_scan_fn bound to the reference's own text of selective_scan_ref
(temp.py:57-139, compiled by oracle/refload.py).  Seeds and shapes are recorded in each file.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def scan_cases(ref_scan):
    """selective_scan_ref on small seeded inputs: grouped / ungrouped B, C; z; D; bias; softplus."""
    out = {}
    cases = [  # name, batch, dim, L, N, G (0 = 3-d B/C), z, D, bias, softplus
        ("k4_full", 2, 24, 37, 16, 4, False, True, True, True),
        ("k4_z", 2, 8, 19, 16, 4, True, True, True, True),
        ("g1_plain", 3, 6, 11, 16, 0, False, False, False, False),
        ("g2_n8", 2, 12, 50, 8, 2, True, True, False, True),
        ("len1", 1, 4, 1, 16, 4, False, True, True, True),
    ]
    for i, (name, b, d, L, N, G, wz, wD, wb, sp) in enumerate(cases):
        g = torch.Generator().manual_seed(100 + i)
        r = lambda *s: torch.randn(*s, generator=g)
        u, dl, A = r(b, d, L), r(b, d, L), -torch.exp(r(d, N))
        Bm = r(b, G, N, L) if G else r(b, N, L)
        Cm = r(b, G, N, L) if G else r(b, N, L)
        D = r(d) if wD else None
        z = r(b, d, L) if wz else None
        bias = r(d) if wb else None
        o, last = ref_scan(u, dl, A, Bm, Cm, D, z, bias, sp, True)
        rec = dict(u=u, delta=dl, A=A, B=Bm, C=Cm, out=o, last_state=last, softplus=int(sp))
        if wD: rec["D"] = D
        if wz: rec["z"] = z
        if wb: rec["delta_bias"] = bias
        out.update({f"{name}.{k}": v for k, v in _np(rec).items()})
    return out


def index_cases(mod):
    """cross-scan / cross-merge of integer-valued grids through the reference's own SS2D code
    (MedMamba.py:256-257, 282-286): selective_scan_fn is replaced by the identity on u."""
    out = {}
    ident = lambda u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False: u
    for (H, W) in [(3, 5), (4, 4), (7, 2), (1, 6)]:
        m = mod.SS2D(d_model=2)           # d_inner 4
        x = torch.arange(2 * 4 * H * W, dtype=torch.float32).view(2, 4, H, W)
        captured = {}

        def spy(u, *a, **kw):
            captured["xs"] = u.clone()
            return u
        mod.selective_scan_fn = spy
        ys = m.forward_corev0(x)
        out[f"{H}x{W}.x"] = x.numpy()
        out[f"{H}x{W}.xs"] = captured["xs"].numpy()
        for i, y in enumerate(ys):
            out[f"{H}x{W}.y{i + 1}"] = y.contiguous().numpy()
    return out


def ss2d_case(mod, ref_scan):
    mod.selective_scan_fn = ref_scan
    out = {}
    for name, d_model, H, W, seed in [("a", 8, 5, 7, 0), ("b", 24, 6, 6, 1), ("c", 48, 9, 4, 2)]:
        torch.manual_seed(seed)
        m = mod.SS2D(d_model=d_model).eval()
        with torch.no_grad():
            # make the SSM term matter: non-integer A, non-unit D, larger x_proj
            m.A_logs.add_(0.3 * torch.randn_like(m.A_logs))
            m.Ds.mul_(0.5).add_(0.2 * torch.randn_like(m.Ds))
            m.x_proj_weight.mul_(4.0)
            m.out_norm.weight.add_(0.1 * torch.randn_like(m.out_norm.weight))
            m.out_norm.bias.add_(0.1 * torch.randn_like(m.out_norm.bias))
            x = torch.randn(2, H, W, d_model)
            y = m(x)
            xz = m.in_proj(x)
            xi = m.act(m.conv2d(xz[..., :m.d_inner].permute(0, 3, 1, 2).contiguous()))
            ys = m.forward_core(xi)
        rec = {f"sd.{k}": v for k, v in m.state_dict().items()}
        rec.update(x=x, y=y, conv_out=xi, y1=ys[0], y2=ys[1], y3=ys[2], y4=ys[3])
        out.update({f"{name}.{k}": v for k, v in _np(rec).items()})
    return out


def vssm_tiny(mod, ref_scan):
    mod.selective_scan_fn = ref_scan
    cfg = dict(depths=[1, 2, 1, 1], dims=[16, 32, 64, 128], num_classes=5)
    torch.manual_seed(3)
    net = mod.VSSM(**cfg).eval()
    with torch.no_grad():
        for m in net.modules():     # non-trivial BatchNorm statistics
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
        x = torch.randn(3, 3, 64, 64)
        logits = net(x)
    rec = {f"sd.{k}": v for k, v in net.state_dict().items()}
    rec.update(x=x, logits=logits, depths=np.array(cfg["depths"]), dims=np.array(cfg["dims"]))
    return _np(rec)


def vssm_t_config1(mod, ref_scan):
    """BASELINE config 1: MedMamba-T, seed 0 weights, batch 8 of 224x224 randn (seed 1), fp32, eval."""
    mod.selective_scan_fn = ref_scan
    torch.manual_seed(0)
    net = mod.VSSM(depths=[2, 2, 4, 2], dims=[96, 192, 384, 768], num_classes=6).eval()
    torch.manual_seed(1)
    x = torch.randn(8, 3, 224, 224)
    with torch.no_grad():
        logits = net(x)
    sd = net.state_dict()
    return _np(dict(logits=logits, weight_seed=0, input_seed=1, head_weight=sd["head.weight"],
                    first_x_proj=sd["layers.0.blocks.0.self_attention.x_proj_weight"]))


def vssm_t_batch256(mod, ref_scan, n=256, chunk=8):
    """The same MedMamba-T (seed 0 weights) on 256 randn images (seed 2), fp32, eval, in chunks of 8 (per-image
    results do not depend on the chunking: eval-mode BatchNorm).  Reference logits for the bf16 top-1 check at one
    full 256-image batch (BASELINE configs[2] lower batch bound)."""
    mod.selective_scan_fn = ref_scan
    torch.manual_seed(0)
    net = mod.VSSM(depths=[2, 2, 4, 2], dims=[96, 192, 384, 768], num_classes=6).eval()
    torch.manual_seed(2)
    x = torch.randn(n, 3, 224, 224)
    outs = []
    with torch.no_grad():
        for i in range(0, n, chunk):
            outs.append(net(x[i:i + chunk]))
            print(f"  b256: {i + chunk}/{n}", flush=True)
    return _np(dict(logits=torch.cat(outs), weight_seed=0, input_seed=2))


SIZES = {   # the sizes train.py:179-182 / test.py:66-72 construct besides 'T'
    "S": dict(depths=[2, 2, 8, 2], dims=[96, 192, 384, 768]),
    "B": dict(depths=[2, 2, 12, 2], dims=[128, 256, 512, 1024]),
    "else": dict(depths=[2, 3, 3, 2], dims=[96, 192, 384, 768]),
}


def vssm_sizes(mod, ref_scan, n=2):
    """MedMamba-S, -B and the default-size branch (train.py:180-182): seed 0 weights, n randn images (seed 3), fp32, eval."""
    mod.selective_scan_fn = ref_scan
    rec = {}
    for name, cfg in SIZES.items():
        torch.manual_seed(0)
        net = mod.VSSM(num_classes=6, **cfg).eval()
        torch.manual_seed(3)
        x = torch.randn(n, 3, 224, 224)
        with torch.no_grad():
            rec[f"{name}.logits"] = net(x)
        rec[f"{name}.depths"], rec[f"{name}.dims"] = np.array(cfg["depths"]), np.array(cfg["dims"])
        print(f"  size {name} done", flush=True)
    rec.update(weight_seed=0, input_seed=3)
    return _np(rec)


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    mod, iface = refload.load_reference()
    ref_scan = iface.selective_scan_ref
    if "--b256" in sys.argv:          # ~6 minutes of CPU; kept apart from the quick fixtures
        np.savez_compressed(os.path.join(OUT, "vssm_t_b256.npz"), **vssm_t_batch256(mod, ref_scan))
        return
    if "--sizes" in sys.argv:         # ~1 minute of CPU
        np.savez_compressed(os.path.join(OUT, "vssm_sizes.npz"), **vssm_sizes(mod, ref_scan))
        return
    np.savez_compressed(os.path.join(OUT, "scan_small.npz"), **scan_cases(ref_scan))
    np.savez_compressed(os.path.join(OUT, "index_maps.npz"), **index_cases(mod))
    np.savez_compressed(os.path.join(OUT, "ss2d_small.npz"), **ss2d_case(mod, ref_scan))
    np.savez_compressed(os.path.join(OUT, "vssm_tiny.npz"), **vssm_tiny(mod, ref_scan))
    np.savez_compressed(os.path.join(OUT, "vssm_t_config1.npz"), **vssm_t_config1(mod, ref_scan))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
