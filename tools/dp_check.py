"""Under torchrun (>= 2 GPUs): the hook-launched, overlapped gradient all-reduce must give the gradients of the plain
after-backward all-reduce bit for bit, with the CNN branch of every block running on a side stream (gradients are then
finalised on two streams).  torchrun --nproc-per-node 2 tools/dp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm
from medmamba_b200 import dist as mdist

rank, world, local = mdist.init_from_env()
dev = torch.device("cuda", local)
torch.manual_seed(0)
net = mm.medmamba_t(6).to(dev).train()
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(32, 3, 224, 224, generator=g).to(dev)
y = torch.randint(0, 6, (32,), generator=g).to(dev)

def grads(overlap):
    red = mdist.GradAllReducer(net.parameters(), overlap=overlap)
    out = []
    for _ in range(3):          # repeated: a race would not hit the same way every time
        net.zero_grad(set_to_none=True)
        torch.manual_seed(1234)     # DropPath draws per step
        red.begin_step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(net(x).float(), y)
        loss.backward()
        red.reduce()
        torch.cuda.synchronize()
        out.append([p.grad.clone() for p in net.parameters()])
    fired = red.launched_in_backward
    red.close()
    return out, fired

plain, _ = grads(False)
over, fired = grads(True)
bad = 0
for run in over + plain[1:]:
    for a, b in zip(run, plain[0]):
        bad += int(not torch.equal(a, b))
t = torch.tensor([bad], device=dev)
torch.distributed.all_reduce(t)
if rank == 0:
    print(f"dp_check: world {world}, buckets launched from hooks {fired}, mismatching gradient tensors {int(t.item())}")
assert t.item() == 0
torch.distributed.destroy_process_group()
