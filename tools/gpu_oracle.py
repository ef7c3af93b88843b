"""GPU oracle: the UNMODIFIED reference (vitok, installed into baseline/_ref by
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`; git-ignored, shipped to the
GPU box by gpurun) run in bf16 on the B200 next to vitok_b200 on the same weights and inputs, at the full sizes of BASELINE
configs 2-4:

    c2  350M-f16x64, 64 x 256 x 256, flash backend          c3  350M-f16x16, 64 mixed-aspect images (128-512 px), sdpa backend
    c4  5B-f16x64, 8 x 512 x 512, flash backend, stress init (gamma, norm weights ~ U(0.5, 1.5): the 44 blocks matter)

Gates: z max-abs <= 6e-2 on valid tokens, PSNR(reconstruction, input) delta <= 0.05 dB.  For context it also prints the
reference's own backend-to-backend noise (flash vs sdpa), and the reference's eager and torch.compile images/s next to ours.

    python tools/gpu_oracle.py [c2 c3 c4] [--no-compile]
"""
import math
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "vitok"))


def import_reference():
    """vitok from baseline/_ref with an empty `webdataset` module (its only missing hard import, vitok/data.py:48)."""
    if "webdataset" not in sys.modules:
        sys.modules["webdataset"] = types.ModuleType("webdataset")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from vitok.models.ae import AE, decode_variant   # noqa: E402
    return AE, decode_variant


def _psnr(a, b):
    mse = float((a.double() - b.double()).pow(2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(4.0 / mse)


def _time(fn, n):
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(workload: str, do_compile: bool = True, timing: bool = True):
    import torch
    import vitok_b200 as vb
    from bench import WORKLOADS, c3_sizes
    RefAE, ref_decode_variant = import_reference()
    variant, B, res, T, backend = WORKLOADS[workload]
    stress = workload == "c4"
    dev = torch.device("cuda", 0)
    cfg = ref_decode_variant(variant)
    assert cfg == vb.decode_variant(variant)
    torch.manual_seed(0)
    with torch.device(dev):
        ref = RefAE(**cfg, attn_backend=backend)
    if stress:
        g = torch.Generator(device=dev).manual_seed(1)
        with torch.no_grad():
            for n, p in ref.named_parameters():
                if "gamma" in n or "norm" in n:
                    p.copy_(torch.rand(p.shape, generator=g, device=dev) + 0.5)
    ref = ref.to(torch.bfloat16).eval()
    ours = vb.AE(**cfg, attn_backend=backend).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.to(dev, torch.bfloat16)
    gi = torch.Generator().manual_seed(1234)
    patch = cfg["spatial_stride"]
    if res == 0:
        sizes = c3_sizes(B, 1234)
        imgs = [(torch.rand(3, h, w, generator=gi) * 2 - 1).to(dev) for h, w in sizes]
    else:
        imgs = (torch.rand(B, 3, res, res, generator=gi) * 2 - 1).to(dev)
    pd = vb.patchify_batch(imgs, patch, T, out_dtype=torch.bfloat16, device=dev)
    mask = pd["patch_mask"]
    valid = mask if backend == "sdpa" else torch.ones_like(mask)
    with torch.no_grad():
        e_r = ref.encode(pd)
        d_r = ref.decode(e_r)
        e_o = ours.encode(pd)
        d_o = ours.decode(e_o)
        # the reference against itself with its other attention backend: how much two bf16 evaluations of the SAME model differ
        other = "sdpa" if backend == "flash" else "flash"
        noise = None
        if not (backend == "sdpa" and not bool(mask.all())):     # flash ignores the mask: only comparable on full batches
            ref.attn_backend = other
            for blk in list(ref.encoder_blocks) + list(ref.decoder_blocks):
                blk.attn.backend = other
            d_n = ref.decode(ref.encode(pd))
            noise = float((d_n["patches"].float() - d_r["patches"].float())[valid].abs().max())
            ref.attn_backend = backend
            for blk in list(ref.encoder_blocks) + list(ref.decoder_blocks):
                blk.attn.backend = backend
    z_err = float((e_o["z"].float() - e_r["z"].float())[valid].abs().max())
    z_rel = float(torch.linalg.norm((e_o["z"].float() - e_r["z"].float())[valid]) / torch.linalg.norm(e_r["z"].float()[valid]))
    p_err = float((d_o["patches"].float() - d_r["patches"].float())[valid].abs().max())
    p_rel = float(torch.linalg.norm((d_o["patches"].float() - d_r["patches"].float())[valid]) / torch.linalg.norm(d_r["patches"].float()[valid]))
    tgt = pd["patches"].float()[valid]
    psnr_r, psnr_o = _psnr(d_r["patches"].float()[valid], tgt), _psnr(d_o["patches"].float()[valid], tgt)
    out = {"workload": workload, "variant": variant, "batch": B, "backend": backend, "stress_init": stress,
           "z_max_abs": z_err, "z_rel_fro": z_rel, "patches_max_abs": p_err, "patches_rel_fro": p_rel,
           "psnr_reference_db": psnr_r, "psnr_ours_db": psnr_o, "psnr_delta_db": abs(psnr_r - psnr_o),
           "reference_backend_noise_patches_max_abs": noise}
    if timing:
        n = 10 if workload != "c4" else 3
        with torch.no_grad():
            t_ours = _time(lambda: ours.decode(ours.encode(pd)), n)
            t_ref = _time(lambda: ref.decode(ref.encode(pd)), n)
            out.update({"ours_images_per_s": B / t_ours * 1e3, "reference_eager_images_per_s": B / t_ref * 1e3})
            if do_compile:
                try:
                    t0 = time.time()
                    enc_c = torch.compile(ref.encode, fullgraph=True)
                    dec_c = torch.compile(ref.decode, fullgraph=True)
                    t_c = _time(lambda: dec_c(enc_c(pd)), n)
                    out.update({"reference_compiled_images_per_s": B / t_c * 1e3, "compile_s": time.time() - t0})
                except Exception as ex:  # noqa: BLE001
                    out["reference_compiled_images_per_s"] = f"torch.compile failed: {type(ex).__name__}: {str(ex)[:200]}"
    del ref, ours
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c3", "c4"]
    if not reference_available():
        print(json.dumps({"unavailable": f"{REF_DIR}/vitok not found (pip install --target baseline/_ref)"}))
        sys.exit(0)
    for w in args:
        r = run(w, do_compile="--no-compile" not in sys.argv)
        print(json.dumps(r), flush=True)
        assert r["z_max_abs"] <= 6e-2, r
        assert r["psnr_delta_db"] <= 0.05, r
