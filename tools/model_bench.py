"""Quick MedMamba-T inference timing (images/s) for development; bench.py is the contract."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--bf16", action="store_true")
args = ap.parse_args()
torch.manual_seed(0)
net = mm.medmamba_t(6).cuda().eval()
if args.unfused:
    for m in net.modules():
        if isinstance(m, mm.SS2D):
            m.fused = False
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
amp = torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.bf16)
with torch.no_grad(), amp:
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        net(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"batch {args.batch} fused={not args.unfused}: {ms:.2f} ms/iter, {args.batch / ms * 1e3:.0f} img/s")
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(x); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
