"""Host side of the reference's training loop around the B200 path.

What the reference's ``train.py`` does per run -- AdamW (+ optional MultiStepLR), a checkpoint dictionary written after
every epoch, ``--resume`` from such a dictionary, one optimisation step per batch -- restated as functions so that a
checkpoint written by the reference resumes here and the other way round:

* ``build_optimizer`` / ``build_scheduler``   train.py:187-200
* ``checkpoint_dict`` / ``save_checkpoint``   train.py:310-319, 352-362 (same keys, same nesting)
* ``resume``                                  train.py:205-249 (same tolerance: every part but the weights is optional)
* ``train_step``                              train.py:277-288, plus the gradient average of the data-parallel config
                                              (``medmamba_b200.dist.GradAllReducer``; the reference is single-GPU)

The datasets, the argparse CLI, logging and early stopping of ``train.py`` are out of scope.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch


def build_optimizer(net: torch.nn.Module, lr: float = 1e-4, npz_dataset: bool = False) -> torch.optim.Optimizer:
    """AdamW as train.py:189-192 builds it: library defaults for the .npz datasets, weight decay 1e-4 otherwise.
    On CUDA parameters the fused multi-tensor implementation of the same update (one kernel per step instead of a
    dozen foreach kernels); its state_dict has the reference's layout, so checkpoints stay interchangeable."""
    fused = all(p.is_cuda for p in net.parameters())
    if npz_dataset:
        return torch.optim.AdamW(net.parameters(), lr=lr, fused=fused)
    return torch.optim.AdamW(net.parameters(), lr=lr, betas=(0.9, 0.999), weight_decay=1e-4, fused=fused)


def build_scheduler(optimizer, milestones: Optional[Sequence[int]]):
    """MultiStepLR(gamma=0.1) when milestones are given (train.py:194-196), else None."""
    if milestones:
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=list(milestones), gamma=0.1)
    return None


def checkpoint_dict(epoch: int, net, optimizer, best_acc: float, num_classes: int, class_indices, scheduler=None) -> dict:
    """The dictionary train.py:310-319 saves; ``net`` may be wrapped (``.module``) by a data-parallel container."""
    model = getattr(net, "module", net)
    d = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
         "best_acc": best_acc, "num_classes": num_classes, "class_indices": class_indices}
    if scheduler is not None:
        d["scheduler_state_dict"] = scheduler.state_dict()
    return d


def save_checkpoint(path: str, *args, **kwargs) -> dict:
    d = checkpoint_dict(*args, **kwargs)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(d, tmp)
    os.replace(tmp, path)           # a killed job never leaves a half-written checkpoint behind
    return d


def resume(path: str, net, optimizer=None, scheduler=None, device=None) -> Tuple[int, float, dict]:
    """Loads a checkpoint in the reference's format.  Returns (start_epoch, best_acc, the raw dictionary).
    Like train.py:208-249, only ``model_state_dict`` is required; a missing optimizer / scheduler / epoch / best_acc
    leaves that part at its fresh value (start_epoch 1, best_acc 0)."""
    ck = torch.load(path, map_location=device if device is not None else "cpu", weights_only=False)
    getattr(net, "module", net).load_state_dict(ck["model_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in ck:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    if scheduler is not None and "scheduler_state_dict" in ck:
        scheduler.load_state_dict(ck["scheduler_state_dict"])
    start_epoch = int(ck["epoch"]) + 1 if "epoch" in ck else 1
    best_acc = float(ck.get("best_acc", 0.0))
    return start_epoch, best_acc, ck


def train_step(net, images: torch.Tensor, labels: torch.Tensor, optimizer, reducer=None,
               autocast_dtype: Optional[torch.dtype] = None, loss_function=None) -> torch.Tensor:
    """One optimisation step (train.py:277-288): zero_grad, forward, cross-entropy, backward, step.  With a
    ``GradAllReducer`` the gradients are averaged over the ranks before the step (overlapped with backward when the
    reducer was built with ``overlap=True``)."""
    loss_function = loss_function or torch.nn.functional.cross_entropy
    optimizer.zero_grad(set_to_none=True)
    if reducer is not None:
        reducer.begin_step()
    with torch.autocast(images.device.type, dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
        logits = net(images)
    loss = loss_function(logits.float(), labels)
    loss.backward()
    if reducer is not None:
        reducer.reduce()
    optimizer.step()
    return loss.detach()
