#!/usr/bin/env python
"""bench.py -- MedMamba-T images/s at 224x224 on N B200 GPUs (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype f32|bf16]
                    [--workload infer|train] [--res R] [--no-extras] [--no-cpu-baseline]

Own arm (default).  A step is one forward pass of MedMamba-T (depths [2,2,4,2], dims [96,192,384,768], 6 classes,
random-init weights, eval mode) over one synthetic batch of B images per GPU -- the batch-sharded inference workload
of BASELINE configs[2].  Every SS2D block runs the hand-written sm_100a kernels (dwconv+SiLU, the 4-direction TMA
scan, out_norm*SiLU(z), shuffle + residual) through the C ABI; linears and the CNN branch are torch (cuBLAS / cuDNN).
  value    : images/s with the input batch resident in HBM, K steps between CUDA events, max over ranks
  e2e      : the same through the public host-facing call (medmamba_b200.InferencePipeline) with HOST (pinned) images:
             every step's H2D copy + forward + logits D2H inside the timed region, the copy of the next batch
             overlapped with the forward of the current one
  roofline : the dominant kernel (ss2d_core_fwd at the stage-1 shape) timed live with CUDA events on its launch
             stream inside the timed steps.  Algorithmic bytes (fused SS2D-core figure with
             per-tensor element sizes): B*L*(D*es_xc + D*es_y + 4*K*(R+2N)) -- xc read once, the merged y written
             once, proj read once.  The kernel is bound by the MUFU exp rate (16/clk/SM), so the HBM fraction is
             small by construction; the fraction of the exp ceiling is reported beside it as "alu".
  roofline_stages : the same two fractions for all four MedMamba-T stage shapes of this run.
  train    : BASELINE configs[3] as a sub-record -- a short data-parallel training step (CE + AdamW, batch 128 per GPU,
             bf16 autocast) with the NCCL gradient all-reduce launched from gradient hooks during backward, timed after
             the inference region; "allreduce_exposed_ms" is the step time with the exchange minus the step time
             without it (N > 1).
  configs  : BASELINE configs[0] (fp32, batch 8), configs[1] (isolated selective_scan_fn at the four stage shapes,
             batch 64, fp32 and bf16, forward and forward+backward) and configs[4] (512x512, batch 32) as sub-records.
  cpu_baseline: the oracle port of the reference's CPU path (oracle.medmamba_ref.vssm_forward + selective_scan_ref)
             on the host cores, bounded sample, rank 0 at N = 1 only.
Multi-GPU (torchrun): one replica per GPU, independent batches, no data-path collective ("weak"); the training
sub-record is the one place with an exchange step.

Reference arm (--impl reference): the reference's own CPU implementation of the path (the oracle port -- the reference
is pure Python and /root/reference does not exist on the GPU box) timed on the host cores for the same metric, 8 images
per step; rank 0 only.
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
N_STATE, K_DIR = 16, 4
MUFU_EXP_PER_S = 16 * 148 * 1.965e9          # MUFU.EX2 ceiling, measured (profiles/README.md)
STAGE_SHAPES = [(384, 3136), (768, 784), (1536, 196), (3072, 49)]      # (K*D, L) of BASELINE configs[1]


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
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(images_per_step: int, steps: int, warmup: int, seed: int = 0, budget_s: float = 600.0):
    """images/s of the oracle port of the reference's CPU path.  Returns (value, per-step images, ms_per_step, cores).
    `images_per_step` is kept unless the whole run would exceed `budget_s` (then it is cut down and reported)."""
    import medmamba_b200 as mm
    from oracle import medmamba_ref
    # every host core: torchrun exports OMP_NUM_THREADS=1, which would leave the CPU arm on a single thread
    ncpu = host_cores()
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    torch.manual_seed(seed)
    sd = {k: v.detach() for k, v in mm.medmamba_t(NUM_CLASSES).state_dict().items()}
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        x1 = torch.randn(1, 3, RES, RES, generator=g)
        medmamba_ref.vssm_forward(sd, x1, depths=tuple(DEPTHS))          # untimed: thread pools, oneDNN primitives
        t0 = time.perf_counter()
        medmamba_ref.vssm_forward(sd, x1, depths=tuple(DEPTHS))
        t_img = time.perf_counter() - t0
        nb = int(max(1, min(images_per_step, budget_s / ((steps + warmup) * t_img))))
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
    value, nb, ms, cores = cpu_reference_rate(8, steps, warmup)
    sample = (f"{nb} images/step x {steps} steps (+{warmup} warm-up), fp32, oracle port of MedMamba.py + "
              f"selective_scan_ref")
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
    if args.workload == "train":
        return (f"MedMamba-T (depths {DEPTHS}, dims {DIMS}) training step (CE + AdamW), batch {args.batch}/GPU, "
                f"{RES}x{RES}x3 synthetic, {NUM_CLASSES} classes (BASELINE configs[3])")
    which = "BASELINE configs[4]" if RES == 512 else ("BASELINE configs[0] shape" if args.dtype == "f32" and args.batch == 8
                                                      else "BASELINE configs[2] shape" if RES == 224 else "off-baseline shape")
    return (f"MedMamba-T (depths {DEPTHS}, dims {DIMS}) inference, batch {args.batch}/GPU, {RES}x{RES}x3 synthetic, "
            f"{NUM_CLASSES} classes ({which})")


# --------------------------------------------------------------------------------------- rooflines
def stage_table(res):
    """(H, D, R) of the four MedMamba-T stages at image side `res`."""
    out = []
    for i, dim in enumerate(DIMS):
        h = res // 4 // (2 ** i)
        out.append((h, dim, -(-(dim // 2) // 16)))
    return out


def core_rooflines(kstats, B, res, dtype, peak):
    """Per stage: algorithmic bytes / exps of one ss2d_core_fwd call over its CUDA-event duration."""
    stages = []
    es_x = 2 if dtype == "bf16" else 4                     # xc element size; proj is fp32
    es_y = 4                                               # the merged y the reference produces is fp32 (MedMamba.py:280)
    for si, (h, D, R) in enumerate(stage_table(res)):
        L = h * h
        key = next((k for k in kstats if k.startswith("ss2d_core_fwd[") and f"B={B},L={L},D={D}," in k), None)
        if key is None:
            continue
        st = kstats[key]
        sec = st["avg_ms"] * 1e-3
        alg_bytes = B * L * (D * es_x + D * es_y + 4 * K_DIR * (R + 2 * N_STATE))
        exps = B * K_DIR * D * L * N_STATE
        stages.append({
            "stage": si + 1, "kernel": key, "avg_ms": round(st["avg_ms"], 4), "launches": st["count"],
            "algorithmic_bytes": alg_bytes, "achieved": round(alg_bytes / sec / 1e9, 1), "unit": "GB/s",
            "frac": round(alg_bytes / sec / 1e9 / peak, 4),
            "interface_equivalent_gbs": round(4 * B * L * (3 * K_DIR * D + 2 * K_DIR * N_STATE) / sec / 1e9, 1),
            "alu": {"bound": "mufu_ex2", "achieved_gexp_s": round(exps / sec / 1e9, 1),
                    "peak_gexp_s": round(MUFU_EXP_PER_S / 1e9, 1), "frac": round(exps / sec / MUFU_EXP_PER_S, 4)}})
    return stages


def timed_loop(fn, steps, dist=None):
    """K calls of fn between CUDA events on the current stream, barrier + synchronize either side -> total ms."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(steps):
        out = fn()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def max_ms(values, dev, dist):
    t = torch.tensor(list(values), device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# --------------------------------------------------------------------------------------- sub-records
def infer_config_record(mm, ops, dev, dist, world, batch, res, dtype, steps, peak, label, graph=False):
    """A short inference measurement of another BASELINE config (fresh model, own warm-up), device-resident inputs."""
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    if dtype == "f32":
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        net = mm.medmamba_t(NUM_CLASSES).to(dev).eval()
        x = torch.randn(batch, 3, res, res, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == "bf16")

        def step():
            with torch.no_grad(), amp:
                return net(x)
        for _ in range(3):
            step()
        timer = ops.KernelTimer()
        ops.set_kernel_timer(timer)
        ms, _ = timed_loop(step, steps, dist)
        ops.set_kernel_timer(None)
        ms = max_ms([ms], dev, dist)[0]
        ks = timer.summary()
        rec = {"workload": label, "dtype": dtype, "batch_per_gpu": batch, "res": res, "steps": steps,
               "ms_per_step": round(ms / steps, 3), "value": round(world * batch * steps / (ms / 1e3), 1), "unit": UNIT,
               "gpu_launches": timer.launches, "roofline_stages": core_rooflines(ks, batch, res, dtype, peak)}
        if graph:
            # the same forward captured once as a CUDA graph (medmamba_b200.GraphedForward, what InferencePipeline uses
            # for small batches): at this size the eager step is bound by the ~250 launches, not by the kernels
            gf = mm.GraphedForward(net, x, torch.bfloat16 if dtype == "bf16" else None)
            for _ in range(3):
                gf.replay()
            msg, _ = timed_loop(gf.replay, steps * 4, dist)
            msg = max_ms([msg], dev, dist)[0]
            rec["cuda_graph"] = {"ms_per_step": round(msg / (steps * 4), 3),
                                 "value": round(world * batch * steps * 4 / (msg / 1e3), 1), "unit": UNIT}
            del gf
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    del net, x
    torch.cuda.empty_cache()
    return rec


def scan_microbench_record(mm, dev, peak, batch=64, iters=10):
    """BASELINE configs[1]: the drop-in selective_scan_fn alone at the four stage shapes, batch 64, fp32 and bf16 I/O,
    forward and forward+backward; GB/s of the interface-level algorithmic bytes es*B*L*(3*KD + 2*K*N) (x3 with the
    backward).  L2 is flushed between calls (a 256 MB write)."""
    import math
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    base_batch = batch
    # the last family repeats the first at 4x the batch: at batch 64 the launch has fewer rows than the machine has lanes
    # (the fused kernel takes the same 0.67 ms at stage 1 there), so that row shows the operator itself
    for dtype, name, layout, batch in ((torch.float32, "f32", "contiguous", base_batch), (torch.float32, "f32", "call_site", base_batch),
                                       (torch.bfloat16, "bf16", "contiguous", base_batch),
                                       (torch.float32, "f32", "contiguous", 4 * base_batch)):
        for KD, L in STAGE_SHAPES:
            g = torch.Generator(device=dev).manual_seed(KD + L)
            rn = lambda *s: torch.randn(*s, device=dev, generator=g)
            u = (0.1 * rn(batch, KD, L)).to(dtype).requires_grad_()
            delta = (0.03 * rn(batch, KD, L)).to(dtype).requires_grad_()
            A = (-torch.arange(1, N_STATE + 1, device=dev, dtype=torch.float32)).repeat(KD, 1).requires_grad_()
            if layout == "contiguous":
                Bm, Cm = (0.05 * rn(batch, K_DIR, N_STATE, L)).requires_grad_(), (0.05 * rn(batch, K_DIR, N_STATE, L)).requires_grad_()
            else:
                # what MedMamba.py:259-261, 267-268 hands over: slices of the x_proj einsum output, N-contiguous with
                # stride(-1) = dt_rank + 2 * d_state
                R = -(-(KD // K_DIR // 2) // 16)
                xdbl = (0.05 * rn(batch, K_DIR, L, R + 2 * N_STATE)).requires_grad_()
                Bm = xdbl[..., R:R + N_STATE].permute(0, 1, 3, 2)
                Cm = xdbl[..., R + N_STATE:].permute(0, 1, 3, 2)
            D = torch.ones(KD, device=dev, requires_grad=True)
            dt = torch.exp(torch.rand(KD, device=dev, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
            bias = (dt + torch.log(-torch.expm1(-dt))).requires_grad_()
            dout = rn(batch, KD, L).to(dtype)
            es = 2 if dtype == torch.bfloat16 else 4
            fwd_bytes = batch * L * (es * 3 * KD + 4 * 2 * K_DIR * N_STATE)

            def fwd():
                with torch.no_grad():
                    return mm.selective_scan_fn(u, delta, A, Bm, Cm, D, None, bias, True)

            def fwd_bwd():
                for t in (u, delta, A, Bm, Cm, D, bias) + ((xdbl,) if layout != "contiguous" else ()):
                    t.grad = None
                mm.selective_scan_fn(u, delta, A, Bm, Cm, D, None, bias, True).backward(dout)

            rec = {"dtype": name, "KD": KD, "L": L, "batch": batch, "bc_layout": layout}
            from medmamba_b200 import ops as _ops
            for label, fn, nbytes in (("fwd", fwd, fwd_bytes), ("fwd_bwd", fwd_bwd, 3 * fwd_bytes)):
                for _ in range(3):
                    fn()
                # kernel time: CUDA events either side of each C-ABI launch (the Python wrapper of a 0.1-1 ms kernel would
                # otherwise show up as idle GPU time between the events)
                timer = _ops.KernelTimer()
                _ops.set_kernel_timer(timer)
                for _ in range(iters):
                    flush.zero_()
                    fn()
                _ops.set_kernel_timer(None)
                ks = timer.summary()
                ms = sum(v["total_ms"] for k, v in ks.items() if k.startswith("scan_")) / iters
                rec[label] = {"ms": round(ms, 4), "gbs": round(nbytes / (ms * 1e-3) / 1e9, 1),
                              "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 4),
                              "kernels": {k.split("[")[0]: round(v["avg_ms"], 4) for k, v in ks.items() if k.startswith("scan_")}}
            out.append(rec)
    del flush
    torch.cuda.empty_cache()
    return out


def train_record(mm, ops, mdist, dev, dist, rank, world, batch, steps, dtype="bf16", kernel_times=False):
    """BASELINE configs[3]: data-parallel training step (CE + AdamW, train.py:187-192, 277-288), batch per GPU, the
    gradient average as NCCL all-reduces over flat buckets launched from gradient hooks during backward."""
    torch.manual_seed(0)                                   # identical replicas
    net = mm.medmamba_t(NUM_CLASSES).to(dev).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)      # train.py:192, one multi-tensor kernel
    red = mdist.GradAllReducer(net.parameters(), overlap=True)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(batch, 3, RES, RES, device=dev, generator=g)
    y = torch.randint(0, NUM_CLASSES, (batch,), device=dev, generator=g)
    amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == "bf16")
    exchange = [True]

    def step():
        opt.zero_grad(set_to_none=True)
        red.begin_step()
        with amp:
            loss = torch.nn.functional.cross_entropy(net(x).float(), y)
        if exchange[0]:
            loss.backward()
            red.finish()
        else:                                   # the same step without the exchange (hooks see a finished bucket list)
            red._works = [True] * len(red.buckets)
            loss.backward()
            red._works = [None] * len(red.buckets)
            red._pending = [len(b) for b in red.buckets]
        opt.step()
        return loss

    for _ in range(3):
        step()
    timer = ops.KernelTimer(timing=kernel_times)
    ops.set_kernel_timer(timer)
    ms, loss = timed_loop(step, steps, dist)
    ops.set_kernel_timer(None)
    fired = red.launched_in_backward
    kstats = timer.summary() if kernel_times else None
    ms_noex = None
    if world > 1:
        exchange[0] = False
        step()
        ms_noex, _ = timed_loop(step, steps, dist)
        exchange[0] = True
    ms, ms_noex = max_ms([ms, ms_noex if ms_noex is not None else 0.0], dev, dist)
    rec = {"workload": f"MedMamba-T training step (CE + AdamW), batch {batch}/GPU, {RES}x{RES} synthetic (BASELINE configs[3])",
           "metric": "MedMamba-T training images/sec at 224x224", "dtype": dtype, "batch_per_gpu": batch,
           "global_batch": world * batch, "steps": steps, "ms_per_step": round(ms / steps, 3),
           "value": round(world * batch * steps / (ms / 1e3), 1), "unit": UNIT, "gpu_launches": timer.launches,
           "parallelism": f"dp{world}, {len(red.buckets)} flat fp32 buckets ({sum(f.numel() for f in red.flat) * 4 / 1e6:.1f} MB), "
                          f"NCCL all-reduce launched from gradient hooks during backward; the CNN branch of every block "
                          f"runs on a side stream, forward and (through autograd) backward",
           "allreduce": None if world == 1 else {
               "buckets": len(red.buckets), "buckets_launched_during_backward": fired,
               "ms_per_step_without_exchange": round(ms_noex / steps, 3),
               "allreduce_exposed_ms": round((ms - ms_noex) / steps, 3),
               "exposed_share_of_step": round((ms - ms_noex) / ms, 4)},
           "loss": round(float(loss.detach()), 4)}
    if kstats:
        rec["kernels"] = {k: {"avg_ms": round(v["avg_ms"], 4), "count": v["count"]} for k, v in sorted(kstats.items())}
    red.close()
    del net, opt, red, x, y
    torch.cuda.empty_cache()
    return rec


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
    from medmamba_b200 import dist as mdist, ops

    torch.backends.cudnn.benchmark = True
    if args.dtype == "f32":
        torch.backends.cudnn.allow_tf32 = False          # every op in true fp32
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = mm.medmamba_t(NUM_CLASSES).to(dev).eval()
    B = args.batch
    x = torch.randn(B, 3, RES, RES, device=dev, generator=torch.Generator(device=dev).manual_seed(1 + rank))
    amp = torch.autocast("cuda", dtype=torch.bfloat16) if args.dtype == "bf16" else torch.autocast("cuda", enabled=False)

    def step():
        with torch.no_grad(), amp:
            return net(x)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM -------------------------------------------
    timer = ops.KernelTimer()
    ops.set_kernel_timer(timer)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, _ = timed_loop(step, args.steps, dist)
    ops.set_kernel_timer(None)
    clocks = sampler.stop() if sampler else None
    launches = timer.launches
    kstats = timer.summary()

    # ---- end to end: the public host-facing call (medmamba_b200.InferencePipeline): every step copies its batch
    # from pinned host memory and reads its logits back; the copy of step i+1 overlaps the forward of step i ----
    pipe = mm.InferencePipeline(net, autocast_dtype=torch.bfloat16 if args.dtype == "bf16" else None)
    x_hosts = [torch.randn(B, 3, RES, RES).pin_memory() for _ in range(2)]
    for _ in pipe.stream(x_hosts):
        pass
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_out = 0
    for logits_host in pipe.stream(x_hosts[i & 1] for i in range(args.steps)):
        n_out += logits_host.shape[0]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert n_out == B * args.steps
    h2d_bytes = x_hosts[0].numel() * 4
    ms_total, e2e_ms = max_ms([ms_total, e2e_s * 1e3], dev, dist)
    del pipe, x_hosts, net, x
    torch.cuda.empty_cache()

    peak, peak_src = measured_peaks()
    extras = {}
    if not args.no_extras:
        tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        extras["train"] = train_record(mm, ops, mdist, dev, dist, rank, world, 128, 5)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        cfgs = {}
        cfgs["configs[0]"] = infer_config_record(mm, ops, dev, dist, world, 8, 224, "f32", 10, peak,
                                                 "MedMamba-T fp32 forward, batch 8, 224x224 (BASELINE configs[0])", graph=True)
        cfgs["configs[4]"] = infer_config_record(mm, ops, dev, dist, world, 32, 512, "bf16", 5, peak,
                                                 "MedMamba-T 512x512 inference, batch 32, stage-1 L = 16384 (BASELINE configs[4])")
        if rank == 0 and world == 1:
            cfgs["configs[1]"] = scan_microbench_record(mm, dev, peak)
        extras["configs"] = cfgs
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "core_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get(f"{args.dtype},{B},{RES}", tj.get(f"{args.dtype},{B}") if RES == 224 else None)
        traffic_src = tj.get("source", "ncu --set full capture of this command, committed under profiles/ (static: not re-measured by this run)")
    stages = core_rooflines(kstats, B, RES, args.dtype, peak)
    roof = None
    if stages:
        s1 = stages[0]
        roof = {"bound": "hbm", "achieved": s1["achieved"], "peak": peak, "unit": "GB/s", "frac": s1["frac"],
                "traffic": traffic, "traffic_source": traffic_src, "kernel": s1["kernel"], "avg_ms": s1["avg_ms"],
                "launches": s1["launches"], "peak_source": peak_src, "algorithmic_bytes": s1["algorithmic_bytes"],
                "interface_equivalent_gbs": s1["interface_equivalent_gbs"], "alu": s1["alu"]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, nb, ms, cores = cpu_reference_rate(8, 2, 0, budget_s=40.0)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 forwards of {nb} images, fp32, oracle port of MedMamba.py + selective_scan_ref ({2 * ms / 1e3:.1f} s)"}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": f"batch-sharded replicas x{world}",
                   "l2": f"inputs ({B * 3 * RES * RES * 4 / 1e6:.0f} MB per step) and activations exceed the 126 MB L2; no flush needed",
                   "batch_sweep": "profiles/README.md (256 / 512 / 1024 per GPU)"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": B * NUM_CLASSES * 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_stages": stages, "cpu_baseline": cpu,
        "kernels": {k: {"avg_ms": round(v["avg_ms"], 4), "count": v["count"]} for k, v in sorted(kstats.items())},
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE configs[3] as the whole run (`--workload train`): one JSON line, same contract."""
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import medmamba_b200 as mm
    from medmamba_b200 import dist as mdist, ops
    mdist.init_from_env("nccl")
    dist = torch.distributed if world > 1 else None
    torch.backends.cudnn.benchmark = True
    sampler = ClockSampler(local) if rank == 0 else None
    rec = train_record(mm, ops, mdist, dev, dist, rank, world, args.batch, args.steps, args.dtype, kernel_times=True)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        print(json.dumps({
            "metric": rec["metric"], "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": 3, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": rec["workload"], "global_batch": rec["global_batch"], "parallelism": rec["parallelism"]},
            "gpu_launches": rec["gpu_launches"], "clocks": clocks, "allreduce": rec["allreduce"], "loss": rec["loss"],
            "kernels": rec.get("kernels")}),
            flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 1024 for inference: the "
                    "top of BASELINE configs[2]'s 256-1024 range; 128 for --workload train)")
    ap.add_argument("--dtype", default="bf16", choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the train / configs sub-records")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"])
    ap.add_argument("--res", type=int, default=224, help="image side; 512 with --batch 32 is BASELINE configs[4]")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 128 if args.workload == "train" else 1024
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
