"""Times the CNN-branch convolutions (MedMamba.py:337-347, eval mode, BN folded) under the cuDNN entry points
torch offers, at the MedMamba-T stage shapes: fused conv+bias+ReLU against plain conv followed by bias+ReLU."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
B = args.batch
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (H, c) in ((56, 48), (28, 96), (14, 192), (7, 384)):
    x = torch.randn(B, c, H, H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for ks in (3, 1):
        w = (torch.randn(c, c, ks, ks, device="cuda", dtype=torch.bfloat16) * 0.05).contiguous(memory_format=torch.channels_last)
        b = torch.randn(c, device="cuda", dtype=torch.bfloat16)
        pad = (ks // 2, ks // 2)
        t_fused = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), pad, (1, 1), 1))
        t_plain = timeit(lambda: F.relu_(F.conv2d(x, w, b, padding=pad)))
        t_nobias = timeit(lambda: F.conv2d(x, w, None, padding=pad))
        gb = 2 * x.numel() * 2 / 1e9
        print(f"H={H} c={c} k={ks}: fused conv+bias+relu {t_fused:.3f} ms | conv(+bias) then relu_ {t_plain:.3f} ms | conv only {t_nobias:.3f} ms | "
              f"in+out {gb:.2f} GB -> {gb / t_fused * 1e3:.0f} GB/s fused")
