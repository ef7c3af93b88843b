"""Pins the CPU oracle (oracle/) against fixtures produced by the running reference.

The fixtures in tests/golden/ come from oracle/make_golden.py, which imports the
unmodified /root/reference.  Bit-exact for pp + variants, <=2e-5 for the fp32 AE.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import ae_oracle, pp_oracle
from oracle.make_golden import PP_CASES, SMALL
from oracle.weights import C3_SIZES, make_state_dict, synth_images


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ppmeta(golden_dir):
    with open(os.path.join(golden_dir, "pp.json")) as f:
        return json.load(f)


def test_decode_variant_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "variants.json")) as f:
        g = json.load(f)
    for v, want in g["ok"].items():
        assert ae_oracle.decode_variant(v) == want, v
    for v, err in g["errors"].items():
        with pytest.raises(Exception) as ei:
            ae_oracle.decode_variant(v)
        assert type(ei.value).__name__ == err


@pytest.mark.parametrize("ci", range(len(PP_CASES)))
def test_patchify_unpatchify_bit_exact(ci, ppmeta, golden_dir):
    H, W, p, T = PP_CASES[ci]
    m = ppmeta[f"case{ci}"]
    img = synth_images([(H, W)], seed=100 + ci)[0]
    d = pp_oracle.patchify(img, p, T)
    for k in ("patches", "patch_mask", "row_idx", "col_idx", "time_idx", "orig_height", "orig_width",
              "grid_rows", "grid_cols"):
        assert sha(np.asarray(d[k])) == m[k], (ci, k)
    canvas = pp_oracle.unpatchify(pp_oracle.collate([d]), patch=p)
    assert list(canvas.shape) == m["unpatchify_shape"]
    assert sha(canvas) == m["unpatchify_sha"]
    assert sha(pp_oracle.convert_format(canvas, "minus_one_to_one", "0_255")) == m["u8_sha"]
    assert sha(pp_oracle.convert_format(canvas, "minus_one_to_one", "zero_to_one")) == m["zero_to_one_sha"]
    # roundtrip exactness (reference tests/cpu/test_pp.py:356-378, tightened to bit-exact)
    assert np.array_equal(canvas[0, :, :H, :W], img)
    small = np.load(os.path.join(golden_dir, "pp_small.npz"))
    if f"case{ci}_patches" in small:
        assert np.array_equal(small[f"case{ci}_patches"], d["patches"])
        assert np.array_equal(small[f"case{ci}_canvas"], canvas)


def test_ragged_batch_bit_exact(ppmeta):
    m = ppmeta["ragged"]
    imgs = synth_images(C3_SIZES[:6], seed=77)
    batch = pp_oracle.collate([pp_oracle.patchify(i, 16, 1024) for i in imgs])
    for k, v in batch.items():
        assert sha(v) == m[k + "_sha"], k
    canvas = pp_oracle.unpatchify(batch, 16)
    assert list(canvas.shape) == m["canvas_shape"] and sha(canvas) == m["canvas_sha"]
    assert sha(pp_oracle.unpatchify(batch, 16, max_grid_size=32)) == m["canvas32_sha"]
    u8 = pp_oracle.convert_format(canvas, "minus_one_to_one", "0_255")
    crops = pp_oracle.unpack(u8, batch["orig_height"], batch["orig_width"])
    assert [list(c.shape) for c in crops] == m["crops_shape"]
    assert [sha(c) for c in crops] == m["crops_sha"]


def _batch(sizes, patch, T, seed):
    b = pp_oracle.collate([pp_oracle.patchify(i, patch, T) for i in synth_images(sizes, seed=seed)])
    return {k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}


@pytest.mark.parametrize("init", ["default", "stress"])
@pytest.mark.parametrize("backend", ["sdpa", "flash"])
def test_ae_small_matches_reference(init, backend, golden_dir):
    g = np.load(os.path.join(golden_dir, "ae_small.npz"))
    cfg = ae_oracle.decode_variant(SMALL)
    sd = make_state_dict(cfg, seed=1 if init == "stress" else 0, stress=(init == "stress"))
    batch = _batch([(128, 128), (96, 64), (50, 120)], 16, 64, seed=5)
    enc = ae_oracle.encode(sd, batch, cfg["encoder_heads"], attn_backend=backend)
    dec = ae_oracle.decode(sd, enc, cfg["decoder_heads"], attn_backend=backend)
    valid = batch["patch_mask"].numpy() if backend == "sdpa" else np.ones_like(batch["patch_mask"].numpy())
    z, p = enc["z"].numpy(), dec["patches"].numpy()
    assert np.abs(z - g[f"{init}_{backend}_z"])[valid].max() <= 2e-5
    assert np.abs(p - g[f"{init}_{backend}_patches"])[valid].max() <= 2e-5


@pytest.mark.parametrize("init", ["default", "stress"])
def test_ae_c1_matches_reference(init, golden_dir):
    """BASELINE.json configs[0]: 350M-f16x64, 4 x 256x256, fp32 on CPU."""
    g = np.load(os.path.join(golden_dir, "ae_c1.npz"))
    cfg = ae_oracle.decode_variant("Ld4-Ld24/1x16x64")
    sd = make_state_dict(cfg, seed=1 if init == "stress" else 0, stress=(init == "stress"))
    batch = _batch([(256, 256)] * 4, 16, 256, seed=1234)
    torch.set_num_threads(os.cpu_count() or 1)
    enc = ae_oracle.encode(sd, batch, cfg["encoder_heads"])
    dec = ae_oracle.decode(sd, enc, cfg["decoder_heads"])
    z, p = enc["z"].numpy(), dec["patches"].numpy()
    assert np.abs(z - g[f"{init}_z"]).max() <= 5e-5
    assert np.abs(p[:, ::4, ::16] - g[f"{init}_patches_sub"]).max() <= 5e-5
    assert np.abs(p.astype(np.float64).sum(-1) - g[f"{init}_patches_rowsum"]).max() <= 2e-3
    # z is per-token zero-mean / unit-variance (LN bottleneck)
    assert np.abs(z.mean(-1)).max() < 1e-5 and np.abs(z.var(-1) - 1).max() < 1e-3


def test_pack_plan_oracle_properties():
    """The token-packing plan (oracle/pp_oracle.py: pack_plan) keeps exactly the keys the reference's mask keeps
    (ae.py:173-187: key j of image b iff patch_mask[b, j]) in their original order, image after image."""
    import numpy as np
    from oracle import pp_oracle
    rng = np.random.RandomState(0)
    mask = rng.rand(9, 300) > 0.3
    mask[4] = False
    mask[5, :129] = True
    for pad, qrows in ((16, 128), (16, 256), (128, 128)):
        pl = pp_oracle.pack_plan(mask, pad, qrows)
        B, N = mask.shape
        assert pl["cu"][0] == 0 and (pl["cu"] % pad == 0).all() and pl["cu"][-1] == len(pl["src"])
        for b in range(B):
            rows = pl["src"][pl["cu"][b]:pl["cu"][b + 1]]
            kept = rows[rows >= 0]
            assert np.array_equal(kept, b * N + np.nonzero(mask[b])[0])          # same keys, same order
            assert (rows[len(kept):] == -1).all() and len(rows) - len(kept) < pad
            g = pl["grp_img"][pl["cuq"][b]:pl["cuq"][b + 1]]
            assert (g == b).all() and len(g) == -(-int(pl["n_valid"][b]) // qrows)   # groups cover the image's valid rows
        valid = pl["rel"] >= 0
        assert np.array_equal(valid.reshape(B, N), mask)
        flat_rows = (pl["cu"][:-1, None] + pl["rel"].reshape(B, N))[mask]
        assert np.array_equal(pl["src"][flat_rows], np.nonzero(mask.reshape(-1))[0])   # rel/cu invert src
        kt = ((pl["n_valid"] + 127) // 128)[pl["grp_img"]][pl["grp_order"]]
        assert (np.diff(kt) <= 0).all() and np.array_equal(np.sort(pl["grp_order"]), np.arange(len(pl["grp_img"])))
