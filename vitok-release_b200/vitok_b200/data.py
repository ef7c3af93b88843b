"""``patch_collate_fn`` -- host-side mirror of vitok/data.py:77-94.

The B200 path never needs it (``patchify_batch`` writes the batched buffers directly), but callers
that collate per-image dicts themselves get the reference behaviour: tensors are stacked, python
scalars become a tensor, anything else stays a list.
"""
from __future__ import annotations

import torch


def patch_collate_fn(batch):
    if not batch:
        return {}
    first = batch[0]
    if isinstance(first, torch.Tensor):
        return torch.stack(batch, dim=0)
    out = {}
    for key in first.keys():
        vals = [item[key] for item in batch]
        if isinstance(vals[0], torch.Tensor):
            out[key] = torch.stack(vals, dim=0)
        elif isinstance(vals[0], (int, float)):
            out[key] = torch.tensor(vals)
        else:
            out[key] = vals
    return out
