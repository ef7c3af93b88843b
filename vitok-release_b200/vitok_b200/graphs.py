"""CUDA-graph capture of the encode / decode hot path.

A full-size batch keeps the GPU busy for ~12 ms per step and the ~90 launches of an encode+decode are hidden behind it, but
at the small per-GPU batches of a strong-scaling or latency-bound deployment (8 images per GPU are
~1.6 ms of GPU work) the host-side launch loop becomes the bottleneck.  Every C-ABI call of this path is capturable
(explicit stream, no allocation, no host sync -- with token packing even the packed row count stays on the device), so the
whole encode -> decode sequence for one (B, N) shape is captured once and replayed with ONE launch.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

_STATIC_KEYS = ("patches", "z", "patch_mask", "row_idx", "col_idx")


class GraphedAE:
    """``GraphedAE(model, example_batch)(batch)`` == ``model.decode(model.encode(batch))`` (or only the half the model has),
    replayed from a CUDA graph.  ``batch`` must have the shapes / dtypes of ``example_batch``; its tensors are copied into
    static buffers, the outputs are static tensors that the next call overwrites (clone them to keep them).  The model must
    be in eval mode; re-create the object after changing weights (``load_state_dict``, ``.to()``, ``quantize()``).
    """

    def __init__(self, model, example_batch: Dict[str, torch.Tensor], warmup: int = 2):
        if model.training:
            raise RuntimeError("GraphedAE: put the model in eval mode first (model.eval())")
        self.model = model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedAE: the model must be on a CUDA device")
        self.static_in = {k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        self._sig = model._signature()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():      # warm-up on a side stream: weight packing, workspaces, allocator
            for _ in range(max(warmup, 1)):
                self._forward(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._forward(self.static_in)

    def _forward(self, d):
        m = self.model
        if m.is_encoder:
            d = m.encode(d)
        if m.is_decoder:
            d = m.decode(d)
        return d

    def __call__(self, batch: Dict[str, torch.Tensor]) -> Dict[str, Optional[torch.Tensor]]:
        if self.model._signature() != self._sig:
            raise RuntimeError("GraphedAE: the model's parameters changed since capture; build a new GraphedAE")
        for k in _STATIC_KEYS:
            src = batch.get(k)
            dst = self.static_in.get(k)
            if isinstance(dst, torch.Tensor):
                if src is None or src.shape != dst.shape:
                    raise ValueError(f"GraphedAE: '{k}' must have shape {tuple(dst.shape)} like the captured batch")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        out = dict(self.static_out)
        for k in ("orig_height", "orig_width"):                # pass-through metadata of THIS batch (ae.py:209-216)
            out[k] = batch.get(k)
        return out


__all__ = ["GraphedAE"]
