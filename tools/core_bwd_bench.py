"""Times the fused SS2D core backward kernel alone at a MedMamba-T stage shape."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medmamba_b200 import ops
from medmamba_b200._lib import lib, ptr, stream_ptr, check

STAGES = [(56, 56, 96, 3), (28, 28, 192, 6), (14, 14, 384, 12), (7, 7, 768, 24)]
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--stage", type=int, default=0)
ap.add_argument("--bf16", action="store_true", help="bf16 xc (the autocast training path)")
args = ap.parse_args()
H, W, D, R = STAGES[args.stage]
B, N = args.batch, 16
g = torch.Generator(device="cuda").manual_seed(0)
xc = 0.1 * torch.randn(B, H, W, D, device="cuda", generator=g)
rp = ops.dt_pad(R)
proj = 0.05 * torch.randn(B, H, W, 4, 32 + rp, device="cuda", generator=g)
Wdt = torch.randn(4, D, R, device="cuda", generator=g) * R ** -0.5
bias = torch.full((4, D), -4.6, device="cuda")
A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
Ds = torch.ones(4 * D, device="cuda")
if args.bf16:
    xc = xc.bfloat16()
ydir, hsave = ops.ss2d_core(xc, proj, Wdt, bias, A, Ds, N, R, save_states=True)
dY = torch.randn(B, H, W, D, device="cuda", generator=g)
ci = ctypes.c_int
tiles = lib().mmb_ss2d_core_bwd_tiles(ci(B), ci(D))
f32 = dict(dtype=torch.float32, device="cuda")
dudir = torch.empty(B, H, W, 4, D, **f32); dproj = torch.empty(tiles, B, H, W, 4, 32 + rp, **f32)
dA = torch.empty(B, 4 * D, N, **f32); dW = torch.empty(B, 4 * D, rp, **f32); dD = torch.empty(B, 4 * D, **f32); db = torch.empty(B, 4 * D, **f32)
def run():
    st = lib().mmb_ss2d_core_bwd(ptr(xc), ptr(proj), ptr(dY), ptr(Wdt), ptr(bias), ptr(A), ptr(Ds), ptr(hsave), ptr(dudir),
                                 ptr(dproj), ptr(dA), ptr(dW), ptr(dD), ptr(db), ci(B), ci(H), ci(W), ci(D), ci(N), ci(R), ci(rp),
                                 ci(1 if args.bf16 else 0), stream_ptr(xc.device))
    check(st, "bwd")
for _ in range(2): run()
torch.cuda.synchronize()
ts = []
for _ in range(args.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps(dict(stage=args.stage + 1, batch=B, bf16=args.bf16, red=os.environ.get('MMB_BWD_RED', '1'), bwd_ms=round(ts[len(ts) // 2], 4))))
