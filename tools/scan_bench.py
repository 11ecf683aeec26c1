"""Times the interface-level selective scan at the MedMamba-T stage shapes (BASELINE config 2).
CUDA events on the current stream; inputs larger than L2 at batch 64 stage 1, and an L2 flush
between timed iterations for the others."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from medmamba_b200.selective_scan_interface import scan_forward
from tests.util import STAGE_SHAPES, make_scan_inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--peak", type=float, default=6550.7)
    ap.add_argument("--bwd", action="store_true", help="also time forward + backward through selective_scan_fn")
    args = ap.parse_args()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for KD, L in STAGE_SHAPES:
        inp = make_scan_inputs("model", args.batch, KD, L, seed=0)
        g = {k: (v.cuda() if v is not None else None) for k, v in inp.items()}
        u, delta = g["u"].to(dt), g["delta"].to(dt)
        Bm, Cm = g["B"].to(dt), g["C"].to(dt)
        run = lambda: scan_forward(u, delta, g["A"], Bm, Cm, g["D"], None, g["delta_bias"], True)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        es = u.element_size()
        nbytes = es * args.batch * L * (3 * KD + 2 * 4 * 16)
        gbs = nbytes / ms / 1e6
        fb = None
        if args.bwd:
            from medmamba_b200 import selective_scan_fn
            leaves = [t.clone().requires_grad_(True) for t in (u, delta, g["A"], Bm, Cm, g["D"], g["delta_bias"])]
            dout = torch.randn_like(u)
            def run_fb():
                for t in leaves:
                    t.grad = None
                out = selective_scan_fn(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], None, leaves[6], True)
                out.backward(dout)
            for _ in range(3):
                run_fb()
            torch.cuda.synchronize()
            tb = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run_fb(); e1.record()
                torch.cuda.synchronize()
                tb.append(e0.elapsed_time(e1))
            tb.sort()
            fb = round(tb[len(tb) // 2], 4)
        print(json.dumps(dict(shape=[args.batch, KD, L], dtype=args.dtype, ms=round(ms, 4), min_ms=round(times[0], 4), fwd_bwd_ms=fb,
                              GBps=round(gbs, 1), frac=round(gbs / args.peak, 3),
                              state_updates_per_ns=round(args.batch * KD * L * 16 / ms / 1e6, 2))))


if __name__ == "__main__":
    main()
