"""Kernel breakdown of one MedMamba-T training step (development aid)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--bf16", action="store_true")
args = ap.parse_args()
torch.manual_seed(0)
net = mm.medmamba_t(6).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
y = torch.randint(0, 6, (args.batch,), device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
        loss = torch.nn.functional.cross_entropy(net(x).float(), y)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=64))
