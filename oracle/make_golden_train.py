"""Generates tests/golden/ref_train_tiny.pt and tests/golden/ref_gradcam_tiny.npz from the UNMODIFIED reference model
(build container only; TEST INFRASTRUCTURE, see oracle/__init__.py).  Run from the repo root:
    python -m oracle.make_golden_train

ref_train_tiny.pt -- a checkpoint in the reference's format (train.py:310-319: epoch, model_state_dict,
    optimizer_state_dict, scheduler_state_dict, best_acc, num_classes, class_indices) written after two AdamW steps
    of the reference model (train.py:187-192, 277-288) on seeded data, plus what the reference itself computes NEXT:
    the logits of the restored model on a probe batch, the loss of the third step and two weights after it.  A resume
    on the B200 path must reproduce those.
ref_gradcam_tiny.npz -- the Grad-CAM consumer's view of the reference (test.py:101-108, grad_cam/utils.py:5-49):
    activations and gradients of ``layers[-1].blocks[-1].conv33conv33conv11[-2]`` for a batch-1 input, backward from
    the top logit, through the reference's own selective_scan_ref under autograd.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CFG = dict(depths=[1, 1], dims=[16, 32], num_classes=3, drop_path_rate=0.0)


def _data(step):
    g = torch.Generator().manual_seed(100 + step)
    return torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 3, (4,), generator=g)


def train_fixture(mod, ref_scan):
    mod.selective_scan_fn = ref_scan
    torch.manual_seed(11)
    net = mod.VSSM(**CFG).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-4)      # train.py:192
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 3], gamma=0.1)                 # train.py:195
    losses = []
    for step in range(2):                                                                           # train.py:277-288
        x, y = _data(step)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    sched.step()
    ck = {"epoch": 1, "model_state_dict": {k: v.clone() for k, v in net.state_dict().items()},
          "optimizer_state_dict": opt.state_dict(), "best_acc": 0.625, "num_classes": 3,
          "class_indices": {0: "a", 1: "b", 2: "c"}, "scheduler_state_dict": sched.state_dict()}     # train.py:310-319
    import copy
    ck = copy.deepcopy(ck)
    # what the reference computes next
    net.eval()
    xp, _ = _data(99)
    with torch.no_grad():
        probe_logits = net(xp).clone()
    net.train()
    x, y = _data(2)
    opt.zero_grad()
    loss3 = torch.nn.functional.cross_entropy(net(x), y)
    loss3.backward()
    opt.step()
    sd = net.state_dict()
    ck["expected"] = {"losses_before": losses, "probe_logits": probe_logits, "loss_step3": float(loss3),
                      "lr_after_resume": opt.param_groups[0]["lr"],
                      "head.weight": sd["head.weight"].clone(),
                      "x_proj_weight": sd["layers.0.blocks.0.self_attention.x_proj_weight"].clone(),
                      "A_logs": sd["layers.0.blocks.0.self_attention.A_logs"].clone(), "cfg": CFG}
    torch.save(ck, os.path.join(OUT, "ref_train_tiny.pt"))


def gradcam_fixture(mod, ref_scan):
    mod.selective_scan_fn = ref_scan
    torch.manual_seed(12)
    net = mod.VSSM(**CFG).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    target = net.layers[-1].blocks[-1].conv33conv33conv11[-2]                                        # test.py:101
    acts, grads = [], []
    h1 = target.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone()))
    h2 = target.register_full_backward_hook(lambda m, gi, go: grads.append(go[0].detach().clone()))  # grad_cam/utils.py:20-23
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    logits = net(x)
    net.zero_grad()
    logits[0, logits.argmax()].backward()                                                            # grad_cam/utils.py:131-161
    h1.remove(); h2.remove()
    rec = {f"sd.{k}": v.detach().numpy() for k, v in net.state_dict().items()}
    rec.update(x=x.numpy(), logits=logits.detach().numpy(), activation=acts[0].numpy(), gradient=grads[0].numpy())
    np.savez_compressed(os.path.join(OUT, "ref_gradcam_tiny.npz"), **rec)


def main():
    mod, iface = refload.load_reference()
    train_fixture(mod, iface.selective_scan_ref)
    gradcam_fixture(mod, iface.selective_scan_ref)
    for f in ("ref_train_tiny.pt", "ref_gradcam_tiny.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
