"""Host-facing inference call: pinned host images in, host logits out, with the host->device copy of the next
batch overlapped with the forward pass of the current one.

The reference's consumers (test.py:60-75, app_streamlit_demo.py:95-110) call ``net(img.to(device))`` and read
the logits back batch by batch; ``InferencePipeline`` is that loop with two device input buffers, a copy
stream and pinned logits buffers, so that PCIe traffic and the forward pass run concurrently.  Outputs are the
same tensors ``net`` would produce (tests/test_infer_gpu.py).
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch


class GraphedForward:
    """``net(x)`` for one fixed input buffer, captured once as a CUDA graph and replayed.

    A MedMamba-T forward is ~250 kernel launches; below a few dozen images per batch the GPU finishes them faster
    than Python can enqueue them (BASELINE configs[0], batch 8: 8.5 ms per step eager, of which the kernels are a
    fraction).  The graph is captured after three eager warm-up passes on a side stream (cuDNN autotuning and every
    lazy allocation happen there), including the block's two-stream branch overlap, which capture records as a
    fork / join.  ``static_input`` must keep its address: copy new images into it, then call ``replay()``."""

    def __init__(self, net: torch.nn.Module, static_input: torch.Tensor, autocast_dtype: Optional[torch.dtype], pool=None):
        self.net, self.x, self.autocast_dtype = net, static_input, autocast_dtype
        self._weights = list(net.parameters()) + list(net.buffers())
        self._sig = self._signature()
        warm = torch.cuda.Stream(static_input.device)
        warm.wait_stream(torch.cuda.current_stream(static_input.device))
        with torch.cuda.stream(warm):
            for _ in range(3):
                self._eager()
        torch.cuda.current_stream(static_input.device).wait_stream(warm)
        torch.cuda.synchronize(static_input.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = self._eager()

    def _eager(self) -> torch.Tensor:
        with torch.no_grad():
            if self.autocast_dtype is None:
                return self.net(self.x)
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                return self.net(self.x)

    def _signature(self):
        return (len(self._weights), sum(t._version for t in self._weights), sum(t.data_ptr() for t in self._weights))

    def stale(self) -> bool:
        """True when a parameter or buffer changed since the capture (optimizer step, load_state_dict, `.data` rebinding,
        `.to()`).  The captured kernels read tensors derived from the weights at capture time (the BN-folded CNN branch,
        ops._derived_params), so a replay would still compute with the old ones: the owner must capture again."""
        return self._signature() != self._sig

    def pool(self):
        return self.graph.pool()

    def replay(self) -> torch.Tensor:
        """Runs the captured forward on the current stream; returns the (static) logits tensor."""
        self.graph.replay()
        return self.out


class InferencePipeline:
    def __init__(self, net: torch.nn.Module, autocast_dtype: Optional[torch.dtype] = torch.bfloat16, device=None,
                 cuda_graph="auto", graph_max_pixels: int = 64 * 224 * 224):
        self.net = net.eval()
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("InferencePipeline runs on CUDA only (there is no CPU path)")
        self.autocast_dtype = autocast_dtype
        self.copy_stream = torch.cuda.Stream(self.device)
        self._in = [None, None]          # device input buffers
        self._out = [None, None]         # pinned host logits
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of slot s finished
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]   # forward that read slot s finished
        self._done = [torch.cuda.Event(), torch.cuda.Event()]       # D2H of slot s finished
        # cuda_graph: True / False / "auto" (batches of at most graph_max_pixels pixels, where launches dominate)
        self.cuda_graph, self.graph_max_pixels = cuda_graph, graph_max_pixels
        self._graphs = [None, None]      # per slot: (shape, dtype, GraphedForward)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            if self.autocast_dtype is None:
                return self.net(x)
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                return self.net(x)

    def _use_graph(self, x: torch.Tensor) -> bool:
        if self.cuda_graph == "auto":
            return x.shape[0] * x.shape[-1] * x.shape[-2] <= self.graph_max_pixels
        return bool(self.cuda_graph)

    def _run(self, slot: int) -> torch.Tensor:
        x = self._in[slot]
        if not self._use_graph(x):
            return self._forward(x)
        ent = self._graphs[slot]
        if ent is None or ent[0] is not x or ent[1].stale():
            other = self._graphs[slot ^ 1]
            pool = other[1].pool() if other is not None else None       # the two slots replay one after the other
            ent = self._graphs[slot] = (x, GraphedForward(self.net, x, self.autocast_dtype, pool=pool))
        return ent[1].replay()

    def _stage(self, slot: int, host: torch.Tensor) -> None:
        buf = self._in[slot]
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = self._in[slot] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._consumed[slot])        # the previous forward on this buffer is over
            buf.copy_(host, non_blocking=True)
            self._ready[slot].record(self.copy_stream)

    def __call__(self, host_images: torch.Tensor) -> torch.Tensor:
        """One batch, synchronously: (B, 3, H, W) host tensor -> (B, num_classes) fp32 host logits."""
        return next(self.stream([host_images]))

    def stream(self, batches: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        """Yields fp32 host logits per batch, in order.  While batch i is in the forward pass, batch i+1 is on
        its way over PCIe and the logits of batch i-1 are being read back."""
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        pending = None                 # (slot, logits shape) whose D2H is in flight
        cur = next(it, None)
        if cur is None:
            return
        slot = 0
        self._stage(slot, cur)
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                self._stage(slot ^ 1, nxt)
            main.wait_event(self._ready[slot])
            logits = self._run(slot).float()
            self._consumed[slot].record(main)
            out = self._out[slot]
            if out is None or out.shape != logits.shape:
                out = self._out[slot] = torch.empty(logits.shape, dtype=torch.float32).pin_memory()
            out.copy_(logits, non_blocking=True)
            self._done[slot].record(main)
            if pending is not None:
                self._done[pending].synchronize()
                yield self._out[pending].clone()
            pending = slot
            cur, slot = nxt, slot ^ 1
        self._done[pending].synchronize()
        yield self._out[pending].clone()
