"""Small shapes through every hand-written kernel family (development aid; written for `compute-sanitizer --tool memcheck`, which this
pool does not allow any more -- it still runs as a plain smoke of the ragged / degenerate shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm
from medmamba_b200 import ops

torch.manual_seed(0)
dev = "cuda"
# fused inference path, fp32 and bf16 autocast, L-parallel (batch 1) and whole sequences, odd grid
for (B, res, dims) in ((1, 64, [32, 64]), (3, 40, [16, 32])):
    net = mm.VSSM(depths=[1, 1], dims=dims, num_classes=3).to(dev).eval()
    x = torch.randn(B, 3, res, res + 8, device=dev)
    with torch.no_grad():
        a = net(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            b = net(x)
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
# forced segment counts on a ragged grid
from tests.test_ss2d_gpu import _core_inputs_gpu
for segs in ("1", "3"):
    os.environ["MMB_CORE_SEGS"] = segs; os.environ["MMB_CORE_S"] = "1"
    xc, proj, Wdt, bias, A, Ds, N = _core_inputs_gpu(2, 13, 37, 40, 3, seed=1)
    ops.ss2d_core(xc, proj, Wdt, bias, A, Ds, N, 3)
    ops.ss2d_core(xc.bfloat16(), proj, Wdt, bias, A, Ds, N, 3)
os.environ.pop("MMB_CORE_SEGS"); os.environ.pop("MMB_CORE_S")
# training step (fused autograd node), fp32 and bf16
net = mm.VSSM(depths=[1, 1], dims=[16, 32], num_classes=3).to(dev).train()
opt = mm.trainer.build_optimizer(net)
x, y = torch.randn(2, 3, 36, 44, device=dev), torch.tensor([0, 2], device=dev)
mm.trainer.train_step(net, x, y, opt)
mm.trainer.train_step(net, x, y, opt, autocast_dtype=torch.bfloat16)
# drop-in operator: strided call-site views, L = 49, with gradient
R, Nn, L = 3, 16, 49
xdbl = torch.randn(2, 4, L, R + 2 * Nn, device=dev, requires_grad=True)
u = torch.randn(2, 4 * 24, L, device=dev, requires_grad=True)
dl = torch.randn(2, 4 * 24, L, device=dev)
A = -torch.rand(4 * 24, Nn, device=dev)
out = mm.selective_scan_fn(u, dl, A, xdbl[..., R:R + Nn].permute(0, 1, 3, 2), xdbl[..., R + Nn:].permute(0, 1, 3, 2),
                           torch.ones(4 * 24, device=dev), None, torch.zeros(4 * 24, device=dev), True)
out.sum().backward()
with torch.no_grad():
    mm.selective_scan_fn(u, dl, A, xdbl[..., R:R + Nn].permute(0, 1, 3, 2), xdbl[..., R + Nn:].permute(0, 1, 3, 2))
# dstate 22: two launches over strided state slices, forward and backward
N2 = 22
A2 = -torch.rand(4 * 24, N2, device=dev, requires_grad=True)
B2 = torch.randn(2, 4, N2, L, device=dev, requires_grad=True)
C2 = torch.randn(2, 4, N2, L, device=dev)
mm.selective_scan_fn(u, dl, A2, B2, C2, torch.ones(4 * 24, device=dev), dl, None, True).sum().backward()
# dwconv rows kernel at one strip / one channel group (the reciprocal bypass)
for (b, h, w, c) in ((1, 7, 3, 40), (2, 5, 9, 4), (1, 1, 1, 8)):
    xx = torch.randn(b, h, w, c, device=dev).bfloat16()
    ops.dwconv3x3_silu(xx, torch.randn(c, 1, 3, 3, device=dev), torch.randn(c, device=dev), out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("sanitize_smoke ok")
