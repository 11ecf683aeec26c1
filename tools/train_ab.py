"""A/B of environment knobs on the MedMamba-T training step in one process (knobs are read per call).
python tools/train_ab.py --variants "default,MMB_TRAIN_BRANCH_OVERLAP=1,MMB_CNN_DENSE=0" """
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import medmamba_b200 as mm

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--variants", default="default")
ap.add_argument("--graph", action="store_true", help="also time the step captured as one CUDA graph (fwd + bwd + AdamW)")
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
y = torch.randint(0, 6, (args.batch,), device="cuda")
for variant in args.variants.split(","):
    for k in [k for k in os.environ if k.startswith("MMB_")]:
        os.environ.pop(k)
    if variant != "default":
        for kv in variant.split("+"):
            k, v = kv.split("=")
            os.environ[k] = v
    torch.manual_seed(0)
    net = mm.medmamba_t(6).cuda().train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(net(x).float(), y)
        loss.backward(); opt.step()
        return loss
    for _ in range(4): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    import time
    enq = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        step()
        enq = min(enq, (time.perf_counter() - t0) * 1e3)          # host time to enqueue one step into an empty queue
    torch.cuda.synchronize()
    print(f"[{variant}] host enqueue {enq:.2f} ms per step", flush=True)
    if args.graph:
        # a fresh model: the gradient accumulators must be created on a non-default stream for the capture
        del net, opt
        torch.manual_seed(0)
        net = mm.medmamba_t(6).cuda().train()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True, capturable=True)
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(3): step()
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                gloss = torch.nn.functional.cross_entropy(net(x).float(), y)
            gloss.backward(); opt.step()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps): g.replay()
        e1.record(); torch.cuda.synchronize()
        msg = e0.elapsed_time(e1) / args.steps
        print(f"[{variant}] CUDA graph: {msg:.2f} ms per step, {args.batch / msg * 1e3:.0f} img/s, loss {gloss.item():.6f}", flush=True)
        del g
    print(f"[{variant}] {ms:.2f} ms per step, {args.batch / ms * 1e3:.0f} img/s, loss after {args.steps + 4} steps {loss.item():.6f}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, opt
