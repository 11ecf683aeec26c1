#!/usr/bin/env python
"""bench.py -- MedMamba-T images/s at 224x224 on N B200 GPUs (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype f32|bf16]

Own arm (default).  A step is one forward pass of MedMamba-T (depths [2,2,4,2], dims
[96,192,384,768], 6 classes, random-init weights, eval mode) over one synthetic batch of B images
per GPU -- the batch-sharded inference workload of BASELINE configs[2].  Every SS2D block runs the
hand-written sm_100a kernels (dwconv+SiLU, the 4-direction TMA scan, out_norm*SiLU(z), shuffle +
residual) through the C ABI; linears and the CNN branch are torch (cuBLAS / cuDNN).
  value   : images/s with the input batch resident in HBM, K steps between CUDA events, max over ranks
  e2e     : the same through the public host-facing call (medmamba_b200.InferencePipeline) with HOST (pinned)
            images: every step's H2D copy + forward + logits D2H inside the timed region, the copy of the next
            batch overlapped with the forward of the current one
  roofline: the dominant kernel (ss2d_core_fwd at the stage-1 shape, L = 3136) timed live with CUDA
            events on its launch stream inside the timed steps; algorithmic bytes are the fused
            SS2D-core figure es*B*L*(2D + K(R+2N)) of SURVEY.md section 8(d).  The kernel is bound by
            the MUFU exp rate (16/clk/SM), so the HBM fraction is small by construction; the
            fraction of the exp ceiling is reported beside it as "alu".
  cpu_baseline: the oracle port of the reference's CPU path (oracle.medmamba_ref.vssm_forward +
            selective_scan_ref) on the host cores, bounded sample, rank 0 at N = 1 only.
Multi-GPU (torchrun): one replica per GPU, independent batches, no data-path collective ("weak").

Reference arm (--impl reference): the reference's own CPU implementation of the path (the oracle
port -- the reference is pure Python and /root/reference does not exist on the GPU box) timed on the
host cores for the same metric; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "MedMamba-T images/sec at 224x224"
UNIT = "images/s"
DEPTHS, DIMS, NUM_CLASSES, RES = [2, 2, 4, 2], [96, 192, 384, 768], 6, 224
MUFU_EXP_PER_S = 16 * 148 * 1.965e9          # MUFU.EX2 ceiling, measured (profiles/README.md)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms in a thread
    (nvidia-smi -lms 200 as a fallback when pynvml is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self.stop_flag = threading.Event()
        self.proc, self.rows, self.nvml = None, [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read, daemon=True)
                self.t.start()
            except OSError:
                self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=1.0)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = [k for k, bit in names.items() if self.reason_bits & bit]
            clk = sorted(self.samples) or [None]
            return {"sm_mhz": clk[len(clk) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        clk = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": clk[len(clk) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "source": "nvidia-smi"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(budget_s: float, steps: int, warmup: int, seed: int = 0):
    """images/s of the oracle port of the reference's CPU path.  Returns (value, per-step images, ms_per_step, cores)."""
    import medmamba_b200 as mm
    from oracle import medmamba_ref
    # every host core: torchrun exports OMP_NUM_THREADS=1, which would leave the CPU arm on a single thread
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    torch.manual_seed(seed)
    sd = {k: v.detach() for k, v in mm.medmamba_t(NUM_CLASSES).state_dict().items()}
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        x1 = torch.randn(1, 3, RES, RES, generator=g)
        t0 = time.perf_counter()
        medmamba_ref.vssm_forward(sd, x1, depths=tuple(DEPTHS))
        t_img = time.perf_counter() - t0
        nb = int(max(1, min(8, budget_s / ((steps + warmup) * t_img))))
        x = torch.randn(nb, 3, RES, RES, generator=g)
        for _ in range(warmup):
            medmamba_ref.vssm_forward(sd, x, depths=tuple(DEPTHS))
        t0 = time.perf_counter()
        for _ in range(steps):
            medmamba_ref.vssm_forward(sd, x, depths=tuple(DEPTHS))
        dt = time.perf_counter() - t0
    return nb * steps / dt, nb, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, nb, ms, cores = cpu_reference_rate(150.0, steps, warmup)
    sample = f"{nb} images/step x {steps} steps (+{warmup} warm-up), fp32, oracle port of MedMamba.py + selective_scan_ref"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_images_per_step": nb},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"MedMamba-T (depths {DEPTHS}, dims {DIMS}) inference, batch {args.batch}/GPU, {RES}x{RES}x3 synthetic, "
            f"{NUM_CLASSES} classes (BASELINE configs[2] shape)")


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import medmamba_b200 as mm
    from medmamba_b200 import ops

    torch.backends.cudnn.benchmark = True
    if args.dtype == "f32":
        torch.backends.cudnn.allow_tf32 = False          # every op in true fp32
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = mm.medmamba_t(NUM_CLASSES).to(dev).eval()
    B = args.batch
    x = torch.randn(B, 3, RES, RES, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
    amp = torch.autocast("cuda", dtype=torch.bfloat16) if args.dtype == "bf16" else torch.autocast("cuda", enabled=False)

    def step(inp):
        with torch.no_grad(), amp:
            return net(inp)

    for _ in range(max(3, args.warmup)):
        out = step(x)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM -------------------------------------------
    timer = ops.KernelTimer()
    ops.set_kernel_timer(timer)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(x)
    e1.record()
    barrier()
    ops.set_kernel_timer(None)
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    launches = timer.launches
    kstats = timer.summary()

    # ---- end to end: the public host-facing call (medmamba_b200.InferencePipeline): every step copies its batch
    # from pinned host memory and reads its logits back; the copy of step i+1 overlaps the forward of step i ----
    pipe = mm.InferencePipeline(net, autocast_dtype=torch.bfloat16 if args.dtype == "bf16" else None)
    x_hosts = [torch.randn(B, 3, RES, RES).pin_memory() for _ in range(2)]
    for _ in pipe.stream(x_hosts):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for logits_host in pipe.stream(x_hosts[i & 1] for i in range(args.steps)):
        n_out += logits_host.shape[0]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert n_out == B * args.steps
    x_host = x_hosts[0]

    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = t.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    peak, peak_src = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "core_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{args.dtype},{B}")
    roof = None
    L1 = (RES // 4) ** 2
    key = next((k for k in kstats if k.startswith("ss2d_core_fwd") and f"L={L1}," in k), None)
    if key:
        st = kstats[key]
        D, R, N, K, L = 96, 3, 16, 4, L1
        es_x = 2 if args.dtype == "bf16" else 4           # xc element size; proj and ydir are fp32
        alg_bytes = B * L * (D * es_x + D * 4 + 4 * K * (R + 2 * N))
        exps = B * K * D * L * N
        ach = alg_bytes / (st["avg_ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "kernel": key, "avg_ms": round(st["avg_ms"], 4), "launches": st["count"],
                "peak_source": peak_src,
                # what the reference-layout operator would have to move for the same work (interface level)
                "interface_equivalent_gbs": round(4 * B * L * (3 * K * D + 2 * K * N) / (st["avg_ms"] * 1e-3) / 1e9, 1),
                "alu": {"bound": "mufu_ex2", "achieved_gexp_s": round(exps / (st["avg_ms"] * 1e-3) / 1e9, 1),
                        "peak_gexp_s": round(MUFU_EXP_PER_S / 1e9, 1),
                        "frac": round(exps / (st["avg_ms"] * 1e-3) / MUFU_EXP_PER_S, 4)}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, nb, ms, cores = cpu_reference_rate(30.0, 2, 0)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 forwards of {nb} images, fp32, oracle port of MedMamba.py + selective_scan_ref ({2 * ms / 1e3:.1f} s)"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": f"batch-sharded replicas x{world}",
                   "l2": f"inputs ({B * 3 * RES * RES * 4 / 1e6:.0f} MB per step) and activations exceed the 126 MB L2; no flush needed",
                   "batch_sweep": "profiles/README.md (256 / 512 / 1024 per GPU)"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": B * NUM_CLASSES * 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "kernels": {k: {"avg_ms": round(v["avg_ms"], 4), "count": v["count"]} for k, v in sorted(kstats.items())},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE configs[3]: data-parallel training step (CE loss + AdamW, train.py:187-192), batch B per GPU,
    gradients averaged with NCCL all-reduces over flat buckets.  One JSON line, same contract."""
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import medmamba_b200 as mm
    from medmamba_b200 import dist as mdist, ops
    mdist.init_from_env("nccl")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)                                   # identical replicas
    net = mm.medmamba_t(NUM_CLASSES).to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    red = mdist.GradAllReducer(net.parameters())
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(B, 3, RES, RES, device=dev, generator=g)
    y = torch.randint(0, NUM_CLASSES, (B,), device=dev, generator=g)
    amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.dtype == "bf16")

    def step():
        opt.zero_grad(set_to_none=True)
        with amp:
            loss = torch.nn.functional.cross_entropy(net(x).float(), y)
        loss.backward()
        red.reduce()
        opt.step()
        return loss

    for _ in range(max(3, args.warmup)):
        step()
    timer = ops.KernelTimer(timing=False)
    ops.set_kernel_timer(timer)
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    clocks = sampler.stop() if sampler else None
    ms = mdist.max_over_ranks([e0.elapsed_time(e1)], device=dev)[0]
    if rank == 0:
        print(json.dumps({
            "metric": "MedMamba-T training images/sec at 224x224", "value": round(world * B * args.steps / (ms / 1e3), 1),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"MedMamba-T training step (CE + AdamW), batch {B}/GPU, {RES}x{RES} synthetic (BASELINE configs[3])",
                       "global_batch": world * B, "parallelism": f"dp{world}, flat-bucket NCCL all-reduce"},
            "gpu_launches": timer.launches, "clocks": clocks, "loss": round(float(loss), 4)}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step (BASELINE configs[2]: 256-1024)")
    ap.add_argument("--dtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"])
    ap.add_argument("--res", type=int, default=224, help="image side; 512 with --batch 32 is BASELINE configs[4]")
    args = ap.parse_args()
    global RES
    RES = args.res
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
