"""Host-side multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo on CPU).

The SS2D path shards by image: replicas with the batch split across ranks and no
data-path collective.  Training adds exactly one exchange per step -- the gradient average -- done here
with a few large flat buckets (14.5 M parameters = 58 MB fp32: two or three NCCL all-reduces).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_from_env(backend: str | None = None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) share of n items for `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values: Iterable[float], device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


class GradAllReducer:
    """Average gradients over ranks through flat fp32 buckets (one all-reduce per bucket).

    Buckets are laid out once in reverse parameter order (the order backward produces gradients in).
    Two ways to drive it:

    * ``reduce()`` after ``loss.backward()``: packs, all-reduces asynchronously, unpacks and scales by 1/world.
    * ``overlap=True``: a post-accumulate hook on every parameter counts the gradients of its bucket; the bucket is
      packed and its all-reduce launched the moment its last gradient is final, so the exchange of the late layers
      travels over NVLink while backward is still computing the early ones (train.py:284 has no counterpart: the
      reference is single-GPU).  ``finish()`` before ``optimizer.step()`` launches whatever did not fire (parameters
      without a gradient contribute zeros so every rank issues identical collectives), waits and unpacks.
    """

    def __init__(self, params, bucket_mb: float = 32.0, overlap: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        cap = int(bucket_mb * (1 << 20) // 4)
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            if cur and size + p.numel() > cap:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]
        self.views: List[List[torch.Tensor]] = []          # per bucket: the slice of the flat buffer of every parameter
        for b, flat in zip(self.buckets, self.flat):
            off, vs = 0, []
            for p in b:
                vs.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.views.append(vs)
        self.overlap = bool(overlap) and self.world > 1
        self._works = [None] * len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        # Streams the hooks of a bucket ran on.  Autograd runs a backward node -- and the gradient accumulation that
        # follows it -- on the stream of its forward; a model that forks a branch onto a side stream (SS_Conv_SSM's CNN
        # branch) therefore finalises some gradients on that stream, and the pack kernel must wait for every one of them.
        self._grad_streams = [set() for _ in self.buckets]
        self._hooks = []
        self.launched_in_backward = 0                       # buckets whose all-reduce started from a hook (last step)
        if self.overlap:
            for bi, bucket in enumerate(self.buckets):
                for p in bucket:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi: int):
        def hook(param):
            if param.is_cuda:
                self._grad_streams[bi].add(torch.cuda.current_stream(param.device))
            self._pending[bi] -= 1
            if self._pending[bi] == 0 and self._works[bi] is None:
                self._launch(bi)
                self.launched_in_backward += 1
        return hook

    def _launch(self, bi: int) -> None:
        bucket, flat, views = self.buckets[bi], self.flat[bi], self.views[bi]
        if flat.is_cuda and self._grad_streams[bi]:
            cur = torch.cuda.current_stream(flat.device)
            for st in self._grad_streams[bi]:
                if st != cur:
                    cur.wait_stream(st)
            self._grad_streams[bi].clear()
        missing = [v for p, v in zip(bucket, views) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        have = [(v, p.grad) for p, v in zip(bucket, views) if p.grad is not None]
        if have:
            # one multi-tensor kernel per bucket instead of one copy per parameter (~300 parameters)
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)

    def _unpack(self, bi: int) -> None:
        bucket, flat, views = self.buckets[bi], self.flat[bi], self.views[bi]
        self._works[bi].wait()
        flat.mul_(1.0 / self.world)
        for p, v in zip(bucket, views):
            if p.grad is None:
                p.grad = v.to(p.dtype).clone()
        torch._foreach_copy_([p.grad for p in bucket], views)

    def finish(self) -> None:
        """Overlapped mode: launch the buckets no hook completed, wait for all of them, write the averages back."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if self._works[bi] is None:
                self._launch(bi)
        for bi in range(len(self.buckets)):
            self._unpack(bi)
        self._works = [None] * len(self.buckets)
        self._pending = [len(b) for b in self.buckets]

    def begin_step(self) -> None:
        """Overlapped mode: call before backward (resets the per-step launch counter)."""
        self.launched_in_backward = 0

    def reduce(self) -> None:
        if self.world == 1:
            return
        if self.overlap:
            return self.finish()
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for bi in range(len(self.buckets)):
            self._unpack(bi)
        self._works = [None] * len(self.buckets)

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
