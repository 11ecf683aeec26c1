"""Times the fused SS2D core kernel alone at the MedMamba-T stage shapes (CUDA events, L2 flushed)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medmamba_b200 import ops

STAGES = [(56, 56, 96, 3), (28, 28, 192, 6), (14, 14, 384, 12), (7, 7, 768, 24)]
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--stage", type=int, default=-1)
ap.add_argument("--noflush", action="store_true")
ap.add_argument("--variants", default="", help="comma list of values of the --env variable to sweep in-process")
ap.add_argument("--env", default="MMB_CORE_CT", help="environment knob swept by --variants (read by the library per call)")
ap.add_argument("--bf16", action="store_true", help="bf16 xc (the autocast layout)")
ap.add_argument("--res", type=int, default=224, help="image side (512 = BASELINE configs[4] stage shapes)")
args = ap.parse_args()
if args.res != 224:
    STAGES = [(args.res // 4 // 2 ** i,) * 2 + (d, r) for i, (d, r) in enumerate(((96, 3), (192, 6), (384, 12), (768, 24)))]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = [(si, v) for si in range(len(STAGES)) for v in (args.variants.split(",") if args.variants else [None])]
for si, variant in cases:
    H, W, D, R = STAGES[si]
    if args.stage >= 0 and si != args.stage:
        continue
    if variant is not None:
        for k in [k for k in os.environ if k.startswith("MMB_CORE_")]:
            os.environ.pop(k)
        if variant != "default":
            if "=" in variant:          # "K1=V1+K2=V2": several knobs at once
                for kv in variant.split("+"):
                    k, v = kv.split("=")
                    os.environ[k] = v
            else:
                os.environ[args.env] = variant
    B, N = args.batch, 16
    g = torch.Generator(device="cuda").manual_seed(0)
    xc = 0.1 * torch.randn(B, H, W, D, device="cuda", generator=g)
    if args.bf16:
        xc = xc.bfloat16()
    rp = ops.dt_pad(R)
    proj = 0.05 * torch.randn(B, H, W, 4, 32 + rp, device="cuda", generator=g)
    Wdt = torch.randn(4, D, R, device="cuda", generator=g) * R ** -0.5
    bias = torch.full((4, D), -4.6, device="cuda")
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
    Ds = torch.ones(4 * D, device="cuda")
    run = lambda: ops.ss2d_core(xc, proj, Wdt, bias, A, Ds, N, R)
    try:
        for _ in range(3):
            run()
    except Exception as e:      # a variant whose ring does not fit this shape
        print(json.dumps(dict(stage=si + 1, batch=B, variant=variant, error=str(e)[:80])))
        continue
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        if not args.noflush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    exps = B * 4 * D * H * W * 16
    fused_bytes = 4 * B * H * W * (2 * D + 4 * (R + 32))
    iface_bytes = 4 * B * H * W * (3 * 4 * D + 2 * 4 * 16)
    plan = ops.core_plan(B, H, W, D, N, R, xc.dtype)
    print(json.dumps(dict(stage=si + 1, batch=B, L=H * W, dtype=str(xc.dtype)[6:], plan=dict(segs=plan[0], ctas_per_sm=plan[1]),
                          ms=round(ms, 4), min_ms=round(ts[0], 4),
                          Gexp_s=round(exps / ms / 1e6, 1), mufu_frac=round(exps / ms / 1e6 / 4653, 3),
                          fused_GBs=round(fused_bytes / ms / 1e6, 1), iface_GBs=round(iface_bytes / ms / 1e6, 1),
                          env={k: v for k, v in os.environ.items() if k.startswith("MMB_")})))
