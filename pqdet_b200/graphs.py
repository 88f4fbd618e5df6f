"""CUDA-graph capture of the training-side hot path (decode + loss forward + backward, all levels).

The fused loss path is two kernel launches per step; at the reference's batch sizes (16 per GPU) the
kernels take ~20 us while eager-mode Python/autograd glue takes ~10x longer.  Capturing the step
into a CUDA graph removes the host from the loop ("CUDA streams and graphs instead of a tracing
compiler").  Inputs are copied into static buffers, the graph is replayed, and the caller reads the
losses and d loss / d head from static outputs.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import config


class GraphedLossStep:
    def __init__(self, head, heads: Sequence[torch.Tensor], target: Sequence[torch.Tensor], warmup: int = 3,
                 autograd: bool = False):
        """head: interpreter.DetectionHead; heads/target: example tensors (shapes are frozen).
        autograd=False (default): the step is head.loss_and_grad -- one kernel in the graph.
        autograd=True: the reference's own call sequence, forward(heads, target)['loss'].mean().backward(),
        captured with its autograd glue kernels (same numbers, a few microseconds more per replay)."""
        self.head = head
        self.autograd = autograd
        self.static_grads = None
        self.static_heads = [h.detach().clone().requires_grad_(True) for h in heads]
        self._sparse = hasattr(target, "tensors")                      # train_dataset.SparseTarget
        self.static_target = target.clone() if self._sparse else tuple(t.detach().clone() for t in target)
        self.graph = torch.cuda.CUDAGraph()
        old = config.nan_check
        config.nan_check = "lazy"                 # no host read inside the captured region
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for h in self.static_heads:
                h.grad = None
            # capture on the stream the warm-up ran on: the library's per-stream workspace is then already
            # allocated and armed (its completion tickets are zero), so the graph holds nothing but the kernels
            with torch.cuda.graph(self.graph, stream=side):
                self.static_out = self._step()
        finally:
            config.nan_check = old
        self.nan_flag = getattr(self.static_out['loss'], 'pq_nan_flag', None)

    def _step(self):
        if not self.autograd:
            out, self.static_grads = self.head.loss_and_grad(self.static_heads, self.static_target)
            return out
        out = self.head(self.static_heads, self.static_target)
        out['loss'].mean().backward()
        return out

    def replay(self):
        """Re-run on whatever is in the static buffers.  -> (loss dict, [d loss / d head])."""
        self.graph.replay()
        if not self.autograd:
            return self.static_out, self.static_grads
        return self.static_out, [h.grad for h in self.static_heads]

    def __call__(self, heads: Sequence[torch.Tensor], target: Sequence[torch.Tensor]):
        for s, h in zip(self.static_heads, heads):
            s.data.copy_(h, non_blocking=True)
        if self._sparse:
            self.static_target.copy_(target)
            return self.replay()
        for s, t in zip(self.static_target, target):
            if s.shape != t.shape:
                raise ValueError("target shape changed: %s vs %s (GT lists must be padded to the captured "
                                 "capacity; use assign_labels(trim=False))" % (tuple(t.shape), tuple(s.shape)))
            s.copy_(t, non_blocking=True)
        return self.replay()

    def check_nan(self):
        """model/loss.py:110-114, once per step on demand (one 4-byte device->host read)."""
        if self.nan_flag is not None and int(self.nan_flag.item()) != 0:
            raise RuntimeError('NaN in loss')
