"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch


def report(name, got, ref, max_abs=None, rel_fro=None):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{name}: non-finite values in output"
    diff = (got - ref).abs()
    ma = diff.max().item()
    rf = (torch.linalg.norm(got - ref) / torch.linalg.norm(ref).clamp_min(1e-30)).item()
    idx = np.unravel_index(int(diff.argmax()), diff.shape)
    print(f"[parity] {name}: max_abs={ma:.4e} rel_fro={rf:.4e} worst@{tuple(int(i) for i in idx)} "
          f"got={got[idx].item():.5f} ref={ref[idx].item():.5f} |ref|max={ref.abs().max().item():.3f}")
    if max_abs is not None:
        assert ma <= max_abs, f"{name}: max_abs {ma} > {max_abs}"
    if rel_fro is not None:
        assert rf <= rel_fro, f"{name}: rel_fro {rf} > {rel_fro}"
    return ma, rf


def bf16_randn(*shape, seed=0, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(device)
