"""Functional CPU restatement of the ViTok-v2 AE encode/decode path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works on a plain state_dict
(reference key names) with torch-CPU tensor ops; dtype follows the weights:
fp32 weights = the exact-math oracle, bf16 weights = the reference's eager
bf16 rounding points, since every op below rounds to the
tensor dtype exactly where the reference's eager op does.

Paths cited are relative to /root/reference.
"""
from __future__ import annotations

import math
import re
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# variant parsing -- vitok/models/ae.py:280-346
# --------------------------------------------------------------------------
_W = {"B": 768, "L": 1024, "G": 1728, "T": 3072, "E": 4096}
_D = {"B": 12, "L": 24, "G": 32, "T": 40, "E": 48}
_H = {"B": 12, "L": 16, "G": 24, "T": 24, "E": 32}
_MLP = 2.67


def parse_variant_name(name: str) -> Dict[str, Any]:
    """vitok/models/ae.py:286-315."""
    if name.startswith("w") and "_d" in name and "_h" in name:
        parts = name.split("_")
        mlp = float(parts[3][1:]) if len(parts) > 3 and parts[3].startswith("m") else _MLP
        return {"width": int(parts[0][1:]), "depth": int(parts[1][1:]), "heads": int(parts[2][1:]), "mlp_factor": mlp}
    wm = re.search(r"w(\d+)", name)
    dm = re.search(r"d(\d+)", name)
    hm = re.search(r"h(\d+)", name)
    mm = re.search(r"m(\d+(?:\.\d+)?)", name)
    base = re.sub(r"w\d+|d\d+|h\d+|m\d+(?:\.\d+)?", "", name)
    if base and base not in _W:
        raise ValueError(f"Unknown base variant: {base}. Available: {list(_W.keys())}")
    return {
        "width": int(wm.group(1)) if wm else _W.get(base, 768),
        "depth": int(dm.group(1)) if dm else _D.get(base, 12),
        "heads": int(hm.group(1)) if hm else _H.get(base, 12),
        "mlp_factor": float(mm.group(1)) if mm else _MLP,
    }


def decode_variant(variant: str) -> Dict[str, Any]:
    """vitok/models/ae.py:318-346."""
    v, rest = variant.split("/")
    enc_v, dec_v = v.split("-") if "-" in v else (v, v)
    parts = list(map(int, rest.split("x")))
    if len(parts) == 3:
        t, s, c = parts
    elif len(parts) == 2:
        t, s, c = 1, parts[0], parts[1]
    else:
        raise ValueError(f"Invalid variant format: {variant}")
    e, d = parse_variant_name(enc_v), parse_variant_name(dec_v)
    return {
        "encoder_width": e["width"], "decoder_width": d["width"],
        "encoder_depth": e["depth"], "decoder_depth": d["depth"],
        "encoder_heads": e["heads"], "decoder_heads": d["heads"],
        "mlp_factor": max(e["mlp_factor"], d["mlp_factor"]),
        "temporal_stride": t, "spatial_stride": s, "channels_per_token": c,
        "pixels_per_token": s * s * t * 3,
    }


def ffn_hidden(width: int, mlp_factor: float) -> int:
    """ae.py:128 ``int(width * mlp_factor)`` then mlp.py:14 round to 16."""
    return ((int(width * mlp_factor) + 8) // 16) * 16


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """vitok/models/modules/norm.py:17-25 -- fp32 compute, cast back to x dtype."""
    x32 = x.float()
    y = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps) * w.float()
    return y.to(x.dtype)


def layer_norm_noaffine(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """vitok/models/modules/norm.py:28-39 -- fp32, biased variance, no affine."""
    x32 = x.float()
    mu = x32.mean(-1, keepdim=True)
    var = (x32 - mu).pow(2).mean(-1, keepdim=True)
    return ((x32 - mu) * torch.rsqrt(var + eps)).to(x.dtype)


def rope_inv_freq(head_dim: int, theta: float = 10000.0) -> torch.Tensor:
    """rotary_embedding.py:7-13 with dim = head_dim // 2 (the per-axis dim, :59)."""
    axis = head_dim // 2
    return 1.0 / (theta ** (torch.arange(0, axis, 2).float() / axis))


def rope_cos_sin(row: torch.Tensor, col: torch.Tensor, head_dim: int, theta: float = 10000.0):
    """rotary_embedding.py:46-75 -- [..., d/2] fp32 cos/sin, row angles then col angles."""
    if head_dim % 4 != 0:
        raise ValueError("2D RoPE requires head dimension divisible by 4")
    inv = rope_inv_freq(head_dim, theta)
    ay = row.to(torch.float32)[..., None] * inv
    ax = col.to(torch.float32)[..., None] * inv
    cos = torch.cat((torch.cos(ay), torch.cos(ax)), dim=-1)
    sin = torch.cat((torch.sin(ay), torch.sin(ax)), dim=-1)
    return cos, sin


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """rotary_embedding.py:102-129 for x [B, N, h, d]; cos/sin [B, N, d/2].

    Interleaved pairs, math in x's dtype with cos/sin cast to it (:118-124).
    Implements the intended [B, N, 1, d/2] broadcast (the reference's shape test
    at :90 mis-broadcasts when N == num_heads).
    """
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    c = cos[:, :, None, :].to(x.dtype)
    s = sin[:, :, None, :].to(x.dtype)
    out_r = xr * c - xi * s
    out_i = xr * s + xi * c
    return torch.stack([out_r, out_i], dim=-1).flatten(3)


def attention_core(q, k, v, key_mask: Optional[torch.Tensor], window: Optional[int] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v, q/k/v [B, N, h, d] -> [B, N, h*d].

    key_mask None  = the flash backend (attention.py:109-117: all N keys).
    key_mask [B,N] = the sdpa backend (attention.py:118-127 + ae.py:173-187):
    mask[b,i,j] = pm[b,i] & pm[b,j]; rows of padded queries are fully masked
    and their output is unspecified in the reference -- zeroed here.
    Softmax in fp32, output cast to q dtype.
    """
    B, N, h, d = q.shape
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) / math.sqrt(d)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    if window is not None and window >= 0:
        # flash_attn_func(window_size=(w, w)) as called at attention.py:113-116: query i sees keys |i - j| <= w.
        # flash-attn is a third-party CUDA-only dependency, so this branch restates its documented semantics
        # (parity unpinned on CPU; tests/test_gpu_attention.py checks it against flash_attn itself on the GPU box).
        idx = torch.arange(N)
        s = s.masked_fill(((idx[:, None] - idx[None, :]).abs() > window)[None, None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    o = torch.matmul(p, vf)
    if key_mask is not None:
        o = o * key_mask[:, None, :, None].float()
    return o.permute(0, 2, 1, 3).reshape(B, N, h * d).to(q.dtype)


def drop_path(x: torch.Tensor, keep: Optional[torch.Tensor], keep_prob: float) -> torch.Tensor:
    """vitok/models/ae.py:15-30 with the random draw made explicit: keep [B] in {0, 1} is floor(keep_prob + U[0,1)) per
    sample (the reference draws it with torch.rand); x / keep_prob * keep.  keep None = eval mode / rate 0: identity."""
    if keep is None:
        return x
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    return x.div(keep_prob) * keep.to(x.dtype).reshape(shape)


def block_forward(sd: Dict[str, torch.Tensor], prefix: str, x, cos, sin, heads: int,
                  key_mask: Optional[torch.Tensor], window: Optional[int] = None,
                  keep: Optional[torch.Tensor] = None, keep_prob: float = 1.0) -> torch.Tensor:
    """vitok/models/ae.py:55-65 (Block.forward), attention.py:92-129, mlp.py:20-23."""
    B, N, D = x.shape
    d = D // heads
    h = rms_norm(x, sd[prefix + "norm1.weight"])
    qkv = F.linear(h, sd[prefix + "attn.qkv_proj.weight"]).reshape(B, N, 3, heads, d)
    q, k, v = qkv.unbind(2)
    q = rms_norm(q, sd[prefix + "attn.norm_q.weight"])
    k = rms_norm(k, sd[prefix + "attn.norm_k.weight"])
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    a = attention_core(q, k, v, key_mask, window)
    attn_out = F.linear(a, sd[prefix + "attn.out_proj.weight"])
    u = F.linear(h, sd[prefix + "ffn.fc1.weight"])
    val, gate = u.chunk(2, dim=-1)
    mlp_out = F.linear(F.silu(gate) * val, sd[prefix + "ffn.fc2.weight"])
    comb = attn_out + mlp_out
    g = sd.get(prefix + "layer_scale.gamma")
    if g is not None:
        comb = comb * g
    return x + drop_path(comb, keep, keep_prob)


def _depth(sd, side: str) -> int:
    n = 0
    while f"{side}_blocks.{n}.norm1.weight" in sd:
        n += 1
    return n


def encode(sd: Dict[str, torch.Tensor], patch_dict: Dict[str, torch.Tensor], heads: int,
           attn_backend: str = "sdpa", theta: float = 10000.0, sw: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """vitok/models/ae.py:189-216.  sw = AE(sw=...) sliding window: flash backend only (attention.py:113-116)."""
    window = sw if (attn_backend == "flash" and sw is not None and sw > 0) else None
    x = F.linear(patch_dict["patches"], sd["patch_embed.weight"], sd["patch_embed.bias"])
    D = x.shape[-1]
    cos, sin = rope_cos_sin(patch_dict["row_idx"], patch_dict["col_idx"], D // heads, theta)
    pm = patch_dict.get("patch_mask")
    key_mask = pm.bool() if (attn_backend == "sdpa" and pm is not None) else None
    for i in range(_depth(sd, "encoder")):
        x = block_forward(sd, f"encoder_blocks.{i}.", x, cos, sin, heads, key_mask, window)
    z = layer_norm_noaffine(F.linear(x, sd["to_code.weight"], sd["to_code.bias"]))
    return {
        "patch_mask": patch_dict.get("patch_mask"), "row_idx": patch_dict["row_idx"],
        "col_idx": patch_dict["col_idx"], "orig_height": patch_dict.get("orig_height"),
        "orig_width": patch_dict.get("orig_width"), "z": z,
    }


def decode(sd: Dict[str, torch.Tensor], enc: Dict[str, torch.Tensor], heads: int,
           attn_backend: str = "sdpa", theta: float = 10000.0, sw: Optional[int] = None,
           drop_keep: Optional[Dict[int, torch.Tensor]] = None, drop_path_rate: float = 0.0) -> Dict[str, torch.Tensor]:
    """vitok/models/ae.py:218-243.  Training-mode stochastic depth (ae.py:143-152: decoder block i has rate
    drop_path_rate * i / (depth - 1)): drop_keep[i] = that block's per-sample 0/1 draw."""
    window = sw if (attn_backend == "flash" and sw is not None and sw > 0) else None
    x = F.linear(enc["z"], sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    D = x.shape[-1]
    cos, sin = rope_cos_sin(enc["row_idx"], enc["col_idx"], D // heads, theta)
    pm = enc.get("patch_mask")
    key_mask = pm.bool() if (attn_backend == "sdpa" and pm is not None) else None
    depth = _depth(sd, "decoder")
    for i in range(depth):
        rate = drop_path_rate * i / max(depth - 1, 1)
        keep = drop_keep.get(i) if (drop_keep is not None and rate > 0.0) else None
        x = block_forward(sd, f"decoder_blocks.{i}.", x, cos, sin, heads, key_mask, window, keep, 1.0 - rate)
    return {
        "patch_mask": enc.get("patch_mask"), "row_idx": enc.get("row_idx"),
        "col_idx": enc.get("col_idx"), "orig_height": enc.get("orig_height"),
        "orig_width": enc.get("orig_width"),
        "patches": F.linear(x, sd["to_pixels.weight"], sd["to_pixels.bias"]),
    }


def psnr(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> float:
    mse = (a.double() - b.double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(data_range * data_range / mse)


def charbonnier_loss(pred: torch.Tensor, target: torch.Tensor, patch_mask: Optional[torch.Tensor], eps: float = 1e-3) -> torch.Tensor:
    """scripts/train_vae.py:314-320 verbatim in meaning: per-token mean over the pixel dim of sqrt(diff^2 + eps^2)
    (diff in fp32), masked, per-image mean over valid tokens (clamp_min 1), batch mean."""
    diff = (pred - target).float()
    per_token = (diff.pow(2) + eps ** 2).sqrt().mean(dim=2)
    if patch_mask is None:
        return per_token.mean(dim=1).mean()
    per_token = per_token * patch_mask.float()
    actual = patch_mask.sum(dim=1).clamp_min(1).float()
    return (per_token.sum(dim=1) / actual).mean()


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 1.0, win: int = 11, sigma: float = 1.5) -> float:
    """Structural similarity of two image batches [B, C, H, W] as torchmetrics' StructuralSimilarityIndexMeasure computes it
    (the reference's reduced-precision gate, tests/gpu/test_float8_inference.py:348-354, vitok/metrics.py): 11 x 11 gaussian
    window (sigma 1.5), K1 = 0.01, K2 = 0.03, per-channel, reflect-padded by half a window and cropped again, mean over
    everything.  Test infrastructure only."""
    a, b = a.double(), b.double()
    C = a.shape[1]
    ax = torch.arange(win, dtype=torch.float64) - (win - 1) / 2
    g1 = torch.exp(-(ax / sigma) ** 2 / 2)
    g1 = g1 / g1.sum()
    kernel = (g1[:, None] * g1[None, :]).expand(C, 1, win, win).contiguous()
    pad = (win - 1) // 2
    ap, bp = F.pad(a, (pad, pad, pad, pad), mode="reflect"), F.pad(b, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat([ap, bp, ap * ap, bp * bp, ap * bp])
    out = F.conv2d(stack, kernel, groups=C)
    mu_a, mu_b, e_aa, e_bb, e_ab = out.split(a.shape[0])
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    va, vb, cab = e_aa - mu_a * mu_a, e_bb - mu_b * mu_b, e_ab - mu_a * mu_b
    s = ((2 * mu_a * mu_b + c1) * (2 * cab + c2)) / ((mu_a * mu_a + mu_b * mu_b + c1) * (va + vb + c2))
    return float(s[..., pad:-pad, pad:-pad].mean())


def train_step_grads(sd: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], enc_heads: int, dec_heads: int,
                     attn_backend: str = "sdpa", eps: float = 1e-3, sw: Optional[int] = None,
                     drop_keep: Optional[Dict[int, torch.Tensor]] = None, drop_path_rate: float = 0.0):
    """The training step of scripts/train_vae.py:304-320,371 (forward, Charbonnier loss, backward) on the CPU oracle
    with torch autograd.  Returns (loss, {name: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    enc = encode(params, batch, enc_heads, attn_backend=attn_backend, sw=sw)
    dec = decode(params, enc, dec_heads, attn_backend=attn_backend, sw=sw, drop_keep=drop_keep, drop_path_rate=drop_path_rate)
    loss = charbonnier_loss(dec["patches"], batch["patches"], batch.get("patch_mask"), eps)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in params.items()}
