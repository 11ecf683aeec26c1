"""Times the bandwidth kernels either side of the scan at the MedMamba-T stage shapes (CUDA events, L2 flushed):
out_norm * SiLU(z) in both slice formats, dwconv3x3 + SiLU, shuffle + residual.  GB/s of the bytes each must move."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medmamba_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--iters", type=int, default=10)
args = ap.parse_args()
B = args.batch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device="cuda").manual_seed(9)
img = torch.randn(B, 3, 224, 224, device="cuda", generator=g)
cw = torch.randn(96, 3, 4, 4, device="cuda", generator=g) * 0.2
cb, lw, lb = torch.randn(96, device="cuda", generator=g), torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
pe_bytes = B * (3 * 224 * 224 * 4 + 56 * 56 * 96 * 4)
for name, mode in (("patch_embed_ln_fp32_fma", False), ("patch_embed_ln_bf16_mma", True)):
    ms = timeit(lambda: ops.patch_embed_ln(img, cw, cb, lw, lb, 1e-5, bf16_math=mode))
    print(json.dumps(dict(stage=0, batch=B, kernel=name, ms=round(ms, 4), GBs=round(pe_bytes / ms / 1e6, 1))))
del img

for si, (H, D) in enumerate(((56, 96), (28, 192), (14, 384), (7, 768))):
    g = torch.Generator(device="cuda").manual_seed(si)
    tok = B * H * H
    xz = torch.randn(B, H, H, 2 * D, device="cuda", generator=g).bfloat16()
    xc = torch.randn(B, H, H, D, device="cuda", generator=g).bfloat16()
    Ds = torch.ones(4 * D, device="cuda")
    gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    y32 = torch.randn(B, H, H, 4, D, device="cuda", generator=g)
    y16 = y32.bfloat16()
    w = torch.randn(D, 1, 3, 3, device="cuda", generator=g)
    bias = torch.randn(D, device="cuda", generator=g)
    left = torch.randn(B, H, H, D // 2, device="cuda", generator=g).bfloat16()
    inp = torch.randn(B, H, H, D, device="cuda", generator=g)
    rows = {
        "outnorm_f32_slices": (lambda: ops.outnorm_gate(y32, xz[..., D:], gamma, beta, 1e-5), tok * D * (16 + 2 + 2)),
        "outnorm_bf16_slices": (lambda: ops.outnorm_gate(y16, xz[..., D:], gamma, beta, 1e-5, xc=xc, Ds=Ds), tok * D * (8 + 2 + 2 + 2)),
        "dwconv_silu_bf16": (lambda: ops.dwconv3x3_silu(xz[..., :D], w, bias, out_dtype=torch.bfloat16), tok * D * 4),
        "shuffle_cat_residual": (lambda: ops.shuffle_cat_residual_raw(left, left, inp), tok * D * (2 + 4 + 4)),
    }
    for name, (fn, nbytes) in rows.items():
        ms = timeit(fn)
        print(json.dumps(dict(stage=si + 1, batch=B, kernel=name, ms=round(ms, 4), GBs=round(nbytes / ms / 1e6, 1))))
