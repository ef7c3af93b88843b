"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only
(the GPU box has no /root/reference):

    python -m oracle.make_golden

The reference needs one stub: ``webdataset`` (vitok/data.py:48 imports it at
module import; it is not installed and not on the hot path).  The AE is built
with attn_backend="sdpa" because flash_attn is CUDA-only;
the flash backend's semantics (no key mask) are obtained by dropping
``patch_mask`` from the input dict, which makes ae.py:199 pass attn_mask=None.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("VITOK_REF", "/root/reference")

sys.path.insert(0, ROOT)
from oracle.weights import make_state_dict, synth_images, C3_SIZES  # noqa: E402


def _import_reference():
    sys.modules.setdefault("webdataset", types.ModuleType("webdataset"))
    sys.path.insert(0, REF)
    import vitok  # noqa: F401
    from vitok.models.ae import AE, decode_variant
    from vitok.pp.ops import patchify, unpatchify, unpack
    from vitok.pp.io import _convert_format, postprocess
    from vitok.data import patch_collate_fn
    return dict(AE=AE, decode_variant=decode_variant, patchify=patchify, unpatchify=unpatchify, unpack=unpack,
                convert=_convert_format, postprocess=postprocess, collate=patch_collate_fn)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


PP_CASES = [  # (H, W, patch, max_tokens)
    (256, 256, 16, 256), (500, 333, 16, 1024), (130, 131, 16, 256), (64, 64, 16, 256), (16, 16, 16, 256),
    (37, 300, 16, 256), (512, 512, 16, 1024), (1, 1, 16, 4), (96, 64, 32, 64), (250, 250, 32, 64),
]

VARIANTS = ["Ld4-Ld24/1x16x16", "Ld4-Ld24/1x16x32", "Ld4-Ld24/1x16x64", "Td4-T/1x16x16", "Td4-T/1x16x32",
            "Td4-T/1x16x64", "Td4-T/1x32x64", "Td4-T/1x32x128", "Td4-T/1x32x256", "Bd2-Bd4/1x16x32", "B/1x16x64",
            "w128_d2_h2-w128_d3_h2/1x16x16", "w256_d1_h4_m3.0-Ld2/16x32", "Lw512h8-Gd3m4/2x8x8", "E/1x16x64"]

SMALL = "w128_d2_h2-w256_d3_h2/1x16x16"   # head_dim 64 (enc) and 128 (dec)


def gen_pp(ref):
    out, meta = {}, {}
    for ci, (H, W, p, T) in enumerate(PP_CASES):
        img = synth_images([(H, W)], seed=100 + ci)[0]
        d = ref["patchify"](p, T)(torch.from_numpy(img))
        rec = {k: v.numpy() for k, v in d.items()}
        meta[f"case{ci}"] = {"H": H, "W": W, "patch": p, "max_tokens": T,
                             **{k: sha(v) for k, v in rec.items()},
                             "scalars": {k: int(rec[k]) for k in ("orig_height", "orig_width", "grid_rows", "grid_cols")}}
        bd = {k: v.unsqueeze(0) for k, v in d.items()}
        canvas = ref["unpatchify"](bd, patch=p)
        meta[f"case{ci}"]["unpatchify_shape"] = list(canvas.shape)
        meta[f"case{ci}"]["unpatchify_sha"] = sha(canvas.numpy())
        u8 = ref["convert"](canvas, "minus_one_to_one", "0_255")
        meta[f"case{ci}"]["u8_sha"] = sha(u8.numpy())
        z2o = ref["convert"](canvas, "minus_one_to_one", "zero_to_one")
        meta[f"case{ci}"]["zero_to_one_sha"] = sha(z2o.numpy())
        if H * W <= 130 * 131:
            out[f"case{ci}_patches"] = rec["patches"]
            out[f"case{ci}_canvas"] = canvas.numpy()
    # batched, ragged (NaFlex) unpatchify incl. max_grid_size and a masked batch
    imgs = synth_images(C3_SIZES[:6], seed=77)
    dicts = [ref["patchify"](16, 1024)(torch.from_numpy(i)) for i in imgs]
    batch = ref["collate"](dicts)
    canvas = ref["unpatchify"](batch, patch=16)
    meta["ragged"] = {"sizes": C3_SIZES[:6], "canvas_shape": list(canvas.shape), "canvas_sha": sha(canvas.numpy()),
                      "canvas32_sha": sha(ref["unpatchify"](batch, patch=16, max_grid_size=32).numpy()),
                      **{k + "_sha": sha(v.numpy()) for k, v in batch.items()}}
    crops = ref["postprocess"](batch, output_format="0_255", do_unpack=True, patch=16)
    meta["ragged"]["crops_sha"] = [sha(c.contiguous().numpy()) for c in crops]
    meta["ragged"]["crops_shape"] = [list(c.shape) for c in crops]
    np.savez_compressed(os.path.join(GOLD, "pp_small.npz"), **out)
    with open(os.path.join(GOLD, "pp.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def gen_variants(ref):
    res = {}
    for v in VARIANTS:
        res[v] = ref["decode_variant"](v)
    bad = {}
    for v in ["Q/1x16x64", "L/16", "L/1x2x3x4"]:
        try:
            ref["decode_variant"](v)
            bad[v] = "ok"
        except Exception as e:  # noqa: BLE001
            bad[v] = type(e).__name__
    with open(os.path.join(GOLD, "variants.json"), "w") as f:
        json.dump({"ok": res, "errors": bad}, f, indent=1, sort_keys=True)


def _naflex_batch(ref, sizes, patch, T, seed):
    imgs = synth_images(sizes, seed=seed)
    return ref["collate"]([ref["patchify"](patch, T)(torch.from_numpy(i)) for i in imgs])


def _run_ae(ref, variant, sd, batch, masked: bool):
    cfg = ref["decode_variant"](variant)
    m = ref["AE"](**cfg, attn_backend="sdpa").eval()
    m.load_state_dict(sd, strict=True)
    d = dict(batch)
    if not masked:
        d.pop("patch_mask")
    with torch.no_grad():
        enc = m.encode(d)
        dec = m.decode(enc)
    return enc["z"].numpy(), dec["patches"].numpy()


def gen_ae_small(ref):
    cfg = ref["decode_variant"](SMALL)
    out = {}
    sizes = [(128, 128), (96, 64), (50, 120)]          # 64, 24, 32 valid tokens of 64
    batch = _naflex_batch(ref, sizes, 16, 64, seed=5)
    for init in ("default", "stress"):
        sd = make_state_dict(cfg, seed=1 if init == "stress" else 0, stress=(init == "stress"))
        for masked in (True, False):
            z, p = _run_ae(ref, SMALL, sd, batch, masked)
            tag = f"{init}_{'sdpa' if masked else 'flash'}"
            out[f"{tag}_z"], out[f"{tag}_patches"] = z, p
    # test_ae.py:43-76 style batch (randn patches, full mask, square grid) on Bd2-Bd4/1x16x32
    np.savez_compressed(os.path.join(GOLD, "ae_small.npz"), **out)


def gen_ae_c1(ref):
    """config 1: 350M-f16x64, 4 x 256x256, fp32 CPU (BASELINE.json configs[0])."""
    variant = "Ld4-Ld24/1x16x64"
    cfg = ref["decode_variant"](variant)
    batch = _naflex_batch(ref, [(256, 256)] * 4, 16, 256, seed=1234)
    out = {}
    for init in ("default", "stress"):
        sd = make_state_dict(cfg, seed=1 if init == "stress" else 0, stress=(init == "stress"))
        z, p = _run_ae(ref, variant, sd, batch, masked=True)
        out[f"{init}_z"] = z.astype(np.float32)
        out[f"{init}_patches_sub"] = p[:, ::4, ::16].astype(np.float32)     # strided sample
        out[f"{init}_patches_rowsum"] = p.astype(np.float64).sum(-1)
        # reference-bf16 self error (calibrates the GPU tolerance)
        m = ref["AE"](**cfg, attn_backend="sdpa").eval()
        m.load_state_dict(sd)
        m = m.to(torch.bfloat16)
        d = {k: (v.to(torch.bfloat16) if v.dtype == torch.float32 else v) for k, v in batch.items()}
        with torch.no_grad():
            e = m.encode(d)
            r = m.decode(e)
        zb, pb = e["z"].float().numpy(), r["patches"].float().numpy()
        out[f"{init}_bf16_z_maxabs"] = np.float64(np.abs(zb - z).max())
        out[f"{init}_bf16_z_relfro"] = np.float64(np.linalg.norm(zb - z) / np.linalg.norm(z))
        out[f"{init}_bf16_p_maxabs"] = np.float64(np.abs(pb - p).max())
        out[f"{init}_bf16_p_relfro"] = np.float64(np.linalg.norm(pb - p) / np.linalg.norm(p))
        print(init, {k: float(v) for k, v in out.items() if k.startswith(init + "_bf16")})
    np.savez_compressed(os.path.join(GOLD, "ae_c1.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = _import_reference()
    gen_variants(ref)
    gen_pp(ref)
    gen_ae_small(ref)
    gen_ae_c1(ref)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
