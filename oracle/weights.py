"""Deterministic weights and inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Independent of any nn.Module
construction order so the same state_dict can be rebuilt on the GPU box (where
/root/reference does not exist) and loaded into the reference AE here.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .ae_oracle import ffn_hidden


def state_dict_shapes(cfg: Dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Key names and shapes of vitok.models.ae.AE.state_dict() (ae.py:115-157)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    P, C = cfg["pixels_per_token"], cfg["channels_per_token"]

    def blocks(side: str, W: int, depth: int, heads: int):
        d = W // heads
        Hf = ffn_hidden(W, cfg["mlp_factor"])
        for i in range(depth):
            p = f"{side}_blocks.{i}."
            out.extend([
                (p + "norm1.weight", (W,)),
                (p + "attn.norm_q.weight", (d,)),
                (p + "attn.norm_k.weight", (d,)),
                (p + "attn.qkv_proj.weight", (3 * W, W)),
                (p + "attn.out_proj.weight", (W, W)),
                (p + "ffn.fc1.weight", (2 * Hf, W)),
                (p + "ffn.fc2.weight", (W, Hf)),
                (p + "layer_scale.gamma", (W,)),
            ])

    if cfg.get("encoder", True):
        We = cfg["encoder_width"]
        out.extend([("patch_embed.weight", (We, P)), ("patch_embed.bias", (We,)),
                    ("to_code.weight", (C, We)), ("to_code.bias", (C,))])
        blocks("encoder", We, cfg["encoder_depth"], cfg["encoder_heads"])
    if cfg.get("decoder", True):
        Wd = cfg["decoder_width"]
        out.extend([("decoder_embed.weight", (Wd, C)), ("decoder_embed.bias", (Wd,)),
                    ("to_pixels.weight", (P, Wd)), ("to_pixels.bias", (P,))])
        blocks("decoder", Wd, cfg["decoder_depth"], cfg["decoder_heads"])
    return out


def make_state_dict(cfg: Dict, seed: int = 0, stress: bool = False, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded weights.

    default: the reference's init statistics: Linear
    weights/biases U(+-1/sqrt(fan_in)), norm weights 1, gamma 1e-4.
    stress : norm weights and gamma ~ U(0.5, 1.5) so the transformer blocks are
    numerically visible (hard part 1).
    Each tensor has its own generator keyed by (seed, crc32(name)).
    """
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in state_dict_shapes(cfg):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith("gamma"):
            t = torch.rand(shape, generator=g) + 0.5 if stress else torch.full(shape, 1e-4)
        elif "norm" in name:
            t = torch.rand(shape, generator=g) + 0.5 if stress else torch.ones(shape)
        else:
            fan_in = shape[1] if len(shape) == 2 else {
                "patch_embed.bias": cfg["pixels_per_token"], "to_code.bias": cfg.get("encoder_width", 0),
                "decoder_embed.bias": cfg["channels_per_token"], "to_pixels.bias": cfg.get("decoder_width", 0)}[name]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.to(dtype)
    return sd


def synth_images(sizes: Sequence[Tuple[int, int]], seed: int = 1234) -> List[np.ndarray]:
    """Uniform [-1, 1] fp32 images [3, H, W] (SURVEY.md section 8d synthetic inputs)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(3, h, w, generator=g) * 2 - 1).numpy() for (h, w) in sizes]


# fixed seeded list for the NaFlex mixed-aspect config (c3): all fit 1024 tokens at p=16
C3_SIZES: List[Tuple[int, int]] = [
    (512, 512), (128, 128), (256, 384), (384, 256), (500, 333), (333, 500), (130, 131), (480, 272),
    (272, 480), (160, 512), (512, 160), (200, 200), (448, 448), (320, 240), (240, 320), (512, 300),
]
