"""numpy restatement of the NaFlex patchify / unpatchify / unpack / format path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def fit_to_token_budget(h: int, w: int, patch: int, max_tokens: int, eps: float = 1e-5) -> Tuple[int, int]:
    """vitok/pp/ops.py:169-196 -- closed-form (h', w') that fits the token budget."""
    h_p = math.ceil(h / patch)
    w_p = math.ceil(w / patch)
    if h_p * w_p <= max_tokens:
        return h, w
    scale = math.sqrt(max_tokens / (h_p * w_p))
    new_h_p = max(1, math.floor(h_p * scale + eps))
    new_w_p = max(1, math.floor(w_p * scale + eps))
    return min(new_h_p * patch, h), min(new_w_p * patch, w)


def patchify(img: np.ndarray, patch: int = 16, max_tokens: int = 256) -> Dict[str, np.ndarray]:
    """vitok/pp/ops.py:217-285.

    img: [C, H, W] float32.  Zero-pads right/bottom to a multiple of ``patch``
    (:235-238), im2col with element order (ch, dy, dx) and token order
    (row-major over the patch grid) (:241-243 -- F.unfold semantics), pads the
    token axis with zeros up to ``max_tokens`` (:259-272) and emits the index
    tensors.  Raises if the grid does not fit (the reference raises a shape
    RuntimeError at :260).
    """
    c, h, w = img.shape
    pad_h = (patch - h % patch) % patch
    pad_w = (patch - w % patch) % patch
    if pad_h or pad_w:
        img = np.pad(img, ((0, 0), (0, pad_h), (0, pad_w)), mode="constant", constant_values=0.0)
    _, hp, wp = img.shape
    gr, gc = hp // patch, wp // patch
    n = gr * gc
    if n > max_tokens:
        raise RuntimeError(f"patch grid {gr}x{gc}={n} exceeds max_tokens={max_tokens}")
    # [C, gr, p, gc, p] -> [gr, gc, C, p, p] -> [n, C*p*p]
    t = img.reshape(c, gr, patch, gc, patch).transpose(1, 3, 0, 2, 4).reshape(n, c * patch * patch)
    patches = np.zeros((max_tokens, c * patch * patch), dtype=img.dtype)
    patches[:n] = t
    mask = np.zeros(max_tokens, dtype=np.bool_)
    mask[:n] = True
    yy, xx = np.meshgrid(np.arange(gr), np.arange(gc), indexing="ij")
    row = np.zeros(max_tokens, dtype=np.int64)
    col = np.zeros(max_tokens, dtype=np.int64)
    row[:n] = yy.reshape(-1)
    col[:n] = xx.reshape(-1)
    return {
        "patches": patches,
        "patch_mask": mask,
        "row_idx": row,
        "col_idx": col,
        "time_idx": np.zeros(max_tokens, dtype=np.int64),
        "orig_height": np.int64(h),
        "orig_width": np.int64(w),
        "grid_rows": np.int64(gr),
        "grid_cols": np.int64(gc),
    }


def collate(dicts: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """vitok/data.py:77-94 -- stack every key along a new batch axis."""
    if not dicts:
        return {}
    return {k: np.stack([np.asarray(d[k]) for d in dicts], axis=0) for k in dicts[0].keys()}


def unpatchify(patch_dict: Dict[str, np.ndarray], patch: int = 16, max_grid_size: Optional[int] = None) -> np.ndarray:
    """vitok/pp/ops.py:295-335.

    Canvas is the batch-wide max(row)+1 x max(col)+1 over valid tokens
    (:319-321) or ``max_grid_size`` squared (:323-324).  Invalid tokens are
    zeroed (:326) and all scatter to the cell their (row, col) names (cell 0
    for patchify output); token 0 is then re-scattered into cell 0 (:332-333).
    Uncovered cells stay 0.  C = 3 is hard-coded (:316).
    """
    patches = np.asarray(patch_dict["patches"])
    mask = np.asarray(patch_dict["patch_mask"]).astype(bool)
    row = np.asarray(patch_dict["row_idx"]).astype(np.int64)
    col = np.asarray(patch_dict["col_idx"]).astype(np.int64)
    B, N, dim = patches.shape
    C = 3
    if max_grid_size is None:
        max_y = int(row[mask].max()) + 1
        max_x = int(col[mask].max()) + 1
    else:
        max_y = max_x = int(max_grid_size)
    p = np.where(mask[..., None], patches, np.zeros((), dtype=patches.dtype))
    flat = row * max_x + col
    tokens = np.zeros((B, max_y * max_x, dim), dtype=patches.dtype)
    for b in range(B):
        # torch.scatter with duplicate indices is order-undefined; patchify never
        # produces duplicates among valid tokens, and invalid tokens write zeros
        # into cell 0 which is then overwritten by token 0 (:332-333).
        tokens[b, flat[b]] = p[b]
        tokens[b, 0] = p[b, 0]
    # F.fold: [B, C*p*p, L] -> [B, C, max_y*p, max_x*p]
    img = tokens.reshape(B, max_y, max_x, C, patch, patch).transpose(0, 3, 1, 4, 2, 5)
    return np.ascontiguousarray(img.reshape(B, C, max_y * patch, max_x * patch))


def unpack(images: np.ndarray, orig_h: Sequence[int], orig_w: Sequence[int]) -> List[np.ndarray]:
    """vitok/pp/ops.py:338-360 -- crop each canvas to its original size."""
    if images.ndim == 3:
        images = images[None]
    return [img[:, : int(h), : int(w)] for img, h, w in zip(images, orig_h, orig_w)]


def convert_format(images: np.ndarray, from_format: str, to_format: str) -> np.ndarray:
    """vitok/pp/io.py:91-121 (fp32 arithmetic, same operation order)."""
    if from_format == to_format:
        return images
    f32 = np.float32
    if to_format == "minus_one_to_one":
        if from_format == "0_255":
            r = images.astype(f32) / f32(127.5) - f32(1.0)
        elif from_format == "zero_to_one":
            r = images * f32(2.0) - f32(1.0)
        else:
            return images
        return np.clip(r, -1.0, 1.0)
    if to_format == "zero_to_one":
        if from_format == "0_255":
            r = images.astype(f32) / f32(255.0)
        elif from_format == "minus_one_to_one":
            r = (images + f32(1.0)) / f32(2.0)
        else:
            return images
        return np.clip(r, 0.0, 1.0)
    if to_format == "0_255":
        if from_format == "minus_one_to_one":
            r = (np.clip(images, -1.0, 1.0) + f32(1.0)) / f32(2.0) * f32(255)
            return np.rint(r).astype(np.uint8)  # torch.round = half-to-even = rint
        if from_format == "zero_to_one":
            return np.rint(np.clip(images, 0.0, 1.0) * f32(255)).astype(np.uint8)
    return images


def normalize_u8(img_u8_hwc: np.ndarray) -> np.ndarray:
    """to_tensor | normalize(minus_one_to_one): vitok/pp/ops.py:140-155.

    torchvision ToTensor: uint8 HWC -> float32 CHW ``/255``; Normalize(0.5, 0.5):
    ``(x - 0.5) / 0.5`` in fp32.
    """
    x = img_u8_hwc.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    return (x - np.float32(0.5)) / np.float32(0.5)


def pack_plan(mask: np.ndarray, pad: int = 16, qrows: int = 128) -> Dict[str, np.ndarray]:
    """Token-packing plan of a [B, N] bool mask: the index layout our packed NaFlex path uses in place of the
    reference's [B,1,N,N] attention mask (vitok/models/ae.py:173-187 keeps key j of image b iff patch_mask[b, j]).

    This is OUR layout, not a reference data structure: the oracle restates its definition in numpy so that the
    CUDA index packing can be checked bit-exactly (include/vitok_b200.h: vtk_pack_plan).  Valid tokens keep their
    order inside an image; every image is padded to a multiple of `pad` packed rows; attention works on groups of
    `qrows` query rows of one image.
    """
    mask = np.asarray(mask).astype(bool)
    B, N = mask.shape
    n_valid = mask.sum(1).astype(np.int32)
    rel = np.where(mask, np.cumsum(mask, axis=1) - 1, -1).astype(np.int32).reshape(-1)
    padded = (n_valid + pad - 1) // pad * pad
    groups = (n_valid + qrows - 1) // qrows
    cu = np.zeros(B + 1, np.int32)
    cu[1:] = np.cumsum(padded)
    cuq = np.zeros(B + 1, np.int32)
    cuq[1:] = np.cumsum(groups)
    src = np.full(int(cu[B]), -1, np.int32)
    grp_img = np.zeros(int(cuq[B]), np.int32)
    for b in range(B):
        idx = np.nonzero(mask[b])[0]
        src[cu[b]:cu[b] + len(idx)] = b * N + idx
        grp_img[cuq[b]:cuq[b + 1]] = b
    # attention work list: groups of images with more 128-key tiles first (any order among equal counts)
    kt = (n_valid + 127) // 128
    grp_order = np.argsort(-kt[grp_img], kind="stable").astype(np.int32)
    return {"n_valid": n_valid, "rel": rel, "cu": cu, "cuq": cuq, "grp_img": grp_img, "grp_order": grp_order, "src": src}
