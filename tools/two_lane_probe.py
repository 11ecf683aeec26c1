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
ap.add_argument("--variants", default="", help="comma list of K=V+K=V environment settings tried in-process (library knobs are read per call)")
ap.add_argument("--stagger", default="", help="comma list of start offsets (ms) for the free-running mode")
ap.add_argument("--priority", action="store_true", help="odd lanes on high-priority streams")
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = mm.medmamba_t(6).cuda().eval()
x = torch.randn(args.batch, 3, 224, 224, device="cuda")
chunks = list(x.chunk(args.lanes))
streams = [torch.cuda.Stream(priority=-1 if (args.priority and i % 2) else 0) for i in range(args.lanes)]

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

def lanes_free(steps, stagger_ms):
    """Each lane runs its own sequence of half-batches with no cross-lane join; lane i starts i * stagger_ms late, so that
    one lane's bandwidth-bound kernels fall into the other lane's scans instead of running in phase with them."""
    main = torch.cuda.current_stream()
    for i, st in enumerate(streams):
        st.wait_stream(main)
        if i and stagger_ms > 0:
            with torch.cuda.stream(st):
                torch.cuda._sleep(int(stagger_ms * 1.9e6 * i))
    out = None
    for _ in range(steps):
        for st, c in zip(streams, chunks):
            with torch.cuda.stream(st), torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(c)
    for st in streams:
        main.wait_stream(st)
    return out


def time_free(steps, stagger_ms):
    lanes_free(2, stagger_ms)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lanes_free(steps, stagger_ms); e1.record(); torch.cuda.synchronize()
    return (e0.elapsed_time(e1) - stagger_ms * (args.lanes - 1)) / steps


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

for variant in (args.variants.split(",") if args.variants else ["default"]):
    for k in [k for k in os.environ if k.startswith("MMB_CORE_") or k.startswith("MMB_PW_")]:
        os.environ.pop(k)
    if variant != "default":
        for kv in variant.split("+"):
            k, v = kv.split("=")
            os.environ[k] = v
    t1, y1 = timeit(one)
    t2, y2 = timeit(lanes)
    if args.stagger:
        for sg in [float(v) for v in args.stagger.split(",")]:
            t3 = time_free(args.steps, sg)
            print(f"[{variant}] free-running lanes, stagger {sg} ms: {t3:.2f} ms per {args.batch} images ({args.batch / t3 * 1e3:.0f} img/s)", flush=True)
    print(f"[{variant}] batch {args.batch}: one stream {t1:.2f} ms ({args.batch / t1 * 1e3:.0f} img/s) | {args.lanes} lanes {t2:.2f} ms "
          f"({args.batch / t2 * 1e3:.0f} img/s) | max |dlogit| {(y1.float() - y2.float()).abs().max().item():.2e}", flush=True)
