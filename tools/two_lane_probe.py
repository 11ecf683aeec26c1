"""Does running two half-batches on two streams beat one full batch on one stream?  (MUFU-bound scan of one
half next to the bandwidth-bound kernels of the other.)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--lanes", type=int, default=2)
ap.add_argument("--steps", type=int, default=8)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = mm.medmamba_t(6).cuda().eval()
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
chunks = list(x.chunk(args.lanes))
streams = [torch.cuda.Stream() for _ in range(args.lanes)]

def one():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return net(x)

def lanes():
    main = torch.cuda.current_stream()
    outs = []
    for st, c in zip(streams, chunks):
        st.wait_stream(main)
        with torch.cuda.stream(st), torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            outs.append(net(c))
    for st in streams:
        main.wait_stream(st)
    return torch.cat(outs)

def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps, y

t1, y1 = timeit(one)
t2, y2 = timeit(lanes)
print(f"batch {args.batch}: one stream {t1:.2f} ms ({args.batch / t1 * 1e3:.0f} img/s) | {args.lanes} lanes {t2:.2f} ms "
      f"({args.batch / t2 * 1e3:.0f} img/s) | max |dlogit| {(y1.float() - y2.float()).abs().max().item():.2e}")
