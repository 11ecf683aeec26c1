"""Kernel breakdown of one MedMamba-T training step, bucketed by kernel family (development aid).
python tools/train_profile.py --batch 128 --bf16"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--rows", type=int, default=40)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = mm.medmamba_t(6).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
y = torch.randint(0, 6, (args.batch,), device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
        loss = torch.nn.functional.cross_entropy(net(x).float(), y)
    loss.backward(); opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
BUCKETS = [("mmb::ss2d_core_bwd", "ours: core bwd"), ("mmb::ss2d_core_fwd", "ours: core fwd"), ("mmb::", "ours: other"),
           ("wgrad", "cuDNN conv wgrad"), ("dgrad", "cuDNN conv dgrad"), ("fprop", "cuDNN conv fprop"), ("conv", "cuDNN conv other"),
           ("nvjet", "GEMM (nvjet)"), ("gemm", "GEMM (cutlass/cublas)"), ("batch_norm", "BatchNorm"), ("bn_", "BatchNorm"),
           ("reduce_kernel", "torch reduce"), ("CatArray", "torch cat"), ("elementwise", "torch elementwise / copy"),
           ("multi_tensor", "optimizer / foreach"), ("FusedAdam", "optimizer / foreach"), ("nchwToNhwc", "layout transforms"),
           ("nhwcToNchw", "layout transforms"), ("softmax", "loss"), ("nll", "loss")]
agg, cnt, rest = collections.Counter(), collections.Counter(), collections.Counter()
for ev in prof.key_averages():
    t = getattr(ev, "self_device_time_total", None) or getattr(ev, "self_cuda_time_total", 0)
    if t <= 0:
        continue
    for key, name in BUCKETS:
        if key in ev.key:
            agg[name] += t; cnt[name] += ev.count
            break
    else:
        agg["other"] += t; cnt["other"] += ev.count; rest[ev.key[:90]] += t
tot = sum(agg.values())
print(f"total device time {tot / 1e3:.2f} ms")
for name, t in agg.most_common():
    print(f"{t / 1e3:8.3f} ms {t / tot * 100:5.1f}%  {cnt[name]:5d} launches  {name}")
print("-- largest unbucketed")
for k, t in rest.most_common(8):
    print(f"{t / 1e3:8.3f} ms  {k}")
if args.rows:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=80))
