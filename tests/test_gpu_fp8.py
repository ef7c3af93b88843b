"""GPU parity: FP8 (e4m3) inference building blocks (reference AE.quantize, vitok/models/ae.py:253-270: torchao
Float8DynamicActivationFloat8WeightConfig = dynamic activation scaling + e4m3 weights on every Linear of the blocks).

torchao is not installed here (third-party, pinned 0.15.0), so the checks restate its arithmetic with torch's
own float8_e4m3fn casts: quantise -> matmul of the dequantised operands in fp32 -> compare.
"""
import pytest
import torch

from _util import bf16_randn, report

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def L():
    from vitok_b200 import _lib
    _lib.load()
    return _lib


@pytest.mark.parametrize("M,K", [(300, 1024), (64, 3760), (1000, 11280), (5, 8)])
def test_quant_rows_e4m3(L, M, K):
    x = bf16_randn(M, K, seed=70, scale=3.0)
    x[0] = 0                                  # an all-zero row: scale 1, zeros
    if M > 2:
        x[2] *= 100.0
    q, scale = L.quant_rows_e4m3(x)
    xf = x.float()
    amax = xf.abs().amax(dim=1)
    ref_scale = torch.where(amax > 0, amax / torch.full_like(amax, 448.0), torch.ones_like(amax))
    assert torch.equal(scale, ref_scale)
    inv = torch.where(amax > 0, torch.full_like(amax, 448.0) / amax, torch.ones_like(amax))   # tensor / tensor = IEEE division
    ref_q = (xf * inv[:, None]).to(torch.float8_e4m3fn)
    assert torch.equal(q.view(torch.uint8), ref_q.view(torch.uint8)), "e4m3 bytes differ from torch's round-to-nearest cast"
    deq = q.float() * scale[:, None]
    assert ((deq - xf).abs() <= 0.0625 * amax[:, None] + 1e-6).all()   # half an ulp at the top binade (3 mantissa bits)


@pytest.mark.parametrize("M,N,K", [(512, 1024, 3760), (300, 256, 944), (4096, 1024, 1024), (130, 3072, 11280)])
def test_fp8_gemm_vs_dequantised_matmul(L, M, N, K):
    a = bf16_randn(M, K, seed=71)
    w = bf16_randn(N, K, seed=72, scale=K ** -0.5)
    a8, a_scale = L.quant_rows_e4m3(a)
    w_scale = float(w.float().abs().max() / 448.0)
    w8 = (w.float() / w_scale).to(torch.float8_e4m3fn)
    gamma = torch.ones(N, dtype=BF, device="cuda")
    x = torch.zeros(M, N, dtype=BF, device="cuda")
    L.proj_residual_fp8(a8, a_scale, w8, w_scale, gamma, x)
    ref = (a8.float() * a_scale[:, None]) @ (w8.float() * w_scale).t()
    report(f"fp8 gemm {M}x{N}x{K} vs dequantised fp32 matmul", x, ref, rel_fro=4e-3)       # only the bf16 rounding of the output
    exact = a.float() @ w.float().t()
    _, rf = report(f"fp8 gemm {M}x{N}x{K} vs the bf16 operands' exact product", x, exact)
    assert rf < 5e-2                                                                      # the quantisation error itself


def _psnr(a, b, peak=2.0):
    mse = ((a.float() - b.float()) ** 2).mean().item()
    return 10.0 * torch.log10(torch.tensor(peak * peak / max(mse, 1e-30))).item()


@pytest.mark.parametrize("backend", ["sdpa", "flash"])
def test_quantized_model_vs_bf16_and_oracle(backend):
    """AE.quantize(): same model in bf16 and in FP8 on a ragged batch under stress init (norm / layer-scale weights ~ U(0.5, 1.5),
    so every block matters).  The reference's acceptance test for its FP8 path is SSIM >= 0.99 of FP8 vs bf16 reconstructions
    (tests/gpu/test_float8_inference.py:348-354); here: relative error of z / patches vs the bf16 run and PSNR between the two
    reconstructions, plus the error against the fp32 CPU oracle next to the bf16 path's own."""
    import numpy as np
    import vitok_b200 as vb
    from oracle import ae_oracle, pp_oracle
    from oracle.weights import make_state_dict, synth_images
    variant = "w256_d2_h4-w512_d3_h4/1x16x16"
    cfg = vb.decode_variant(variant)
    sd = make_state_dict(ae_oracle.decode_variant(variant), seed=4, stress=True)
    model = vb.AE(**cfg, attn_backend=backend).eval()
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda", BF)
    b = pp_oracle.collate([pp_oracle.patchify(i, 16, 256) for i in synth_images([(256, 256), (96, 200), (130, 131), (240, 160)], seed=12)])
    batch = {k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}
    cb = {k: (v.cuda().to(BF) if v.dtype == torch.float32 else v.cuda()) for k, v in batch.items()}
    valid = batch["patch_mask"] if backend == "sdpa" else torch.ones_like(batch["patch_mask"])
    with torch.no_grad():
        e16 = model.encode(cb); d16 = model.decode(e16)
        launches16 = model.last_launch_count
        assert model.quantize() is model and model.quantize() is model      # idempotent, returns self (ae.py:253-270)
        e8 = model.encode(cb); d8 = model.decode(e8)
    assert model.last_launch_count == launches16 + 2 * cfg["decoder_depth"]   # two row-quantisation launches per block
    z16, z8 = e16["z"].cpu().float()[valid], e8["z"].cpu().float()[valid]
    p16, p8 = d16["patches"].cpu().float()[valid], d8["patches"].cpu().float()[valid]
    _, rz = report(f"fp8 vs bf16 z ({backend})", z8, z16)
    _, rp = report(f"fp8 vs bf16 patches ({backend})", p8, p16)
    psnr = _psnr(p8, p16)
    e_o = ae_oracle.encode(sd, batch, cfg["encoder_heads"], attn_backend=backend)
    d_o = ae_oracle.decode(sd, e_o, cfg["decoder_heads"], attn_backend=backend)
    _, r16 = report("bf16 patches vs fp32 oracle", p16, d_o["patches"][valid])
    _, r8 = report("fp8 patches vs fp32 oracle", p8, d_o["patches"][valid])
    print(f"[parity] fp8 vs bf16 reconstruction PSNR {psnr:.1f} dB; rel-Fro vs fp32 oracle: bf16 {r16:.3e}, fp8 {r8:.3e}")
    # e4m3 has 3 mantissa bits: one GEMM on quantised operands is ~3.7 % off (test above); 5 blocks under stress init stay below 10 %
    assert rz < 1e-1 and rp < 1.2e-1 and psnr > 30.0
    assert torch.isfinite(d8["patches"]).all()


def test_per_tensor_quantiser_matches_torch_cast(L):
    """vtk_quant_tensor_e4m3 (torchao's default PerTensor granularity): one scale = amax(tensor) / 448 for every row, bytes equal to
    torch's float8_e4m3fn cast of x * (448 / amax)."""
    x = bf16_randn(777, 1032, seed=50, scale=3.0)
    q, scale = L.quant_tensor_e4m3(x)
    amax = x.float().abs().max()
    assert torch.equal(scale, torch.full_like(scale, float(amax) / 448.0))
    ref_q = (x.float() * (448.0 / amax)).to(torch.float8_e4m3fn)
    assert torch.equal(q.view(torch.uint8), ref_q.view(torch.uint8))


@pytest.mark.parametrize("granularity", ["row", "tensor"])
def test_reference_acceptance_gate_ssim(granularity):
    """The reference's own gate for its reduced-precision path (tests/gpu/test_float8_inference.py:286-354): reconstructions of the
    quantised and of the bf16 model, reshaped to images, must have SSIM >= 0.99 (torchmetrics StructuralSimilarityIndexMeasure,
    data_range = 2.0) and contain no NaN / Inf.  Model: 350M-f16x64 on 4 x 256 x 256 images (BASELINE configs[0]); weights: the
    reference's default random init (pretrained weights need the Hub) -- and, for information, the stress init of the parity tests,
    where every block's output is ~1e4 times larger than at init.  SSIM is oracle/ae_oracle.ssim (torchmetrics is not installed)."""
    import numpy as np
    import vitok_b200 as vb
    from oracle import ae_oracle, pp_oracle
    from oracle.weights import make_state_dict, synth_images
    variant = "Ld4-Ld24/1x16x64"
    cfg = vb.decode_variant(variant)
    b = pp_oracle.collate([pp_oracle.patchify(i, 16, 256) for i in synth_images([(256, 256)] * 4, seed=1234)])
    batch = {k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}
    cb = {k: (v.cuda().to(BF) if v.dtype == torch.float32 else v.cuda()) for k, v in batch.items()}
    scores = {}
    for init in ("default", "stress"):
        sd = make_state_dict(ae_oracle.decode_variant(variant), seed=1 if init == "stress" else 0, stress=init == "stress")
        model = vb.AE(**cfg, attn_backend="flash").eval()
        model.load_state_dict(sd, strict=True)
        model = model.to("cuda", BF)
        model.fp8_activation_scale = granularity
        with torch.no_grad():
            img16 = vb.unpatchify(model.decode(model.encode(cb)), 16, max_grid_size=16).float().cpu()
            model.quantize()
            out8 = model.decode(model.encode(cb))
            img8 = vb.unpatchify(out8, 16, max_grid_size=16).float().cpu()
        assert torch.isfinite(img8).all()
        scores[init] = ae_oracle.ssim(img8, img16, data_range=2.0)
    print(f"[parity] FP8 ({granularity} activation scale) vs bf16 reconstruction SSIM: default init {scores['default']:.5f}, stress init {scores['stress']:.5f}")
    assert scores["default"] >= 0.99
    assert scores["stress"] >= 0.90
