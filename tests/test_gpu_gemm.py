"""GPU parity: tcgen05 GEMM + fused epilogues vs the CPU oracle (oracle/ae_oracle.py), through the C ABI."""
import math

import pytest
import torch
import torch.nn.functional as F

from _util import bf16_randn, report
from oracle import ae_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from vitok_b200 import _lib
    _lib.load()
    return _lib


@pytest.mark.parametrize("N,K", [(128, 128), (256, 64), (64, 256)])
def test_probe_kmajor(L, N, K):
    a = bf16_randn(128, K, seed=1)
    b = bf16_randn(N, K, seed=2)
    d = L.umma_probe(a, b, N, K, False, 0, 0, 0)
    report(f"probe K-major N={N} K={K}", d, a.float().cpu() @ b.float().cpu().T, max_abs=1e-3 * math.sqrt(K) * 4)


@pytest.mark.parametrize("N,K", [(64, 128), (128, 128), (128, 64)])
def test_probe_mn_major(L, N, K):
    """B given as [K, N] row-major (the layout of V in attention): MN-major SWIZZLE_128B descriptor."""
    a = bf16_randn(128, K, seed=3)
    b = bf16_randn(K, N, seed=4)
    ref = a.float().cpu() @ b.float().cpu()
    # descriptor fields for a [K x 64]-blocked, 128B-swizzled MN-major operand: LBO = stride between the
    # 64-column blocks, SBO = stride between 8-row (k) groups, +2048 B per UMMA_K (16 rows).  (Other
    # encodings fault with an illegal shared-memory access, so only the documented one is exercised.)
    d = L.umma_probe(a, b, N, K, True, K * 128, 1024, 2048)
    report(f"probe MN-major N={N} K={K}", d, ref, max_abs=1e-3 * math.sqrt(K) * 4)


@pytest.mark.parametrize("M,N,K,bias", [
    (256, 256, 128, False), (128, 64, 64, True), (300, 776, 200, True), (1024, 1024, 768, True),
    (16384, 1024, 768, True), (512, 768, 1024, True), (384, 1024, 16, True), (2048, 1024, 3760, False),
])
def test_linear(L, M, N, K, bias):
    a = bf16_randn(M, K, seed=10)
    w = bf16_randn(N, K, seed=11, scale=1.0 / math.sqrt(K))
    b = bf16_randn(N, seed=12) if bias else None
    out = L.linear(a, w, b)
    ref = F.linear(a.float().cpu(), w.float().cpu(), b.float().cpu() if bias else None)
    report(f"linear M={M} N={N} K={K}", out, ref, max_abs=4e-2, rel_fro=4e-3)


@pytest.mark.parametrize("M,C,K", [(256, 16, 128), (1000, 64, 1024), (512, 256, 3072), (130, 32, 128), (256, 128, 256)])
def test_linear_ln(L, M, C, K):
    a = bf16_randn(M, K, seed=20)
    w = bf16_randn(C, K, seed=21, scale=1.0 / math.sqrt(K))
    b = bf16_randn(C, seed=22, scale=0.1)
    out = L.linear_ln(a, w, b)
    lin = F.linear(a.float().cpu(), w.float().cpu(), b.float().cpu()).to(torch.bfloat16)
    ref = ae_oracle.layer_norm_noaffine(lin)
    report(f"linear_ln M={M} C={C} K={K}", out, ref.float(), max_abs=8e-2, rel_fro=1e-2)


def _block_inputs(M, D, heads, mlp=2.67, seed=0):
    d = D // heads
    Hf = ae_oracle.ffn_hidden(D, mlp)
    h = bf16_randn(M, D, seed=seed)
    wqkv = bf16_randn(3 * D, D, seed=seed + 1, scale=1 / math.sqrt(D))
    w1 = bf16_randn(2 * Hf, D, seed=seed + 2, scale=1 / math.sqrt(D))
    g = torch.Generator().manual_seed(seed + 3)
    nq = (torch.rand(d, generator=g) + 0.5).to(torch.bfloat16).cuda()
    nk = (torch.rand(d, generator=g) + 0.5).to(torch.bfloat16).cuda()
    side = int(math.ceil(math.sqrt(M)))
    idx = torch.arange(M)
    row, col = (idx // side).cuda(), (idx % side).cuda()
    return d, Hf, h, wqkv, w1, nq, nk, row, col


@pytest.mark.parametrize("M,D,heads", [(256, 128, 2), (1024, 1024, 16), (500, 256, 2), (4096, 1024, 16), (640, 3072, 24)])
def test_qkv_swiglu(L, M, D, heads):
    from vitok_b200.models.ae import pack_w_in
    d, Hf, h, wqkv, w1, nq, nk, row, col = _block_inputs(M, D, heads)
    inv = ae_oracle.rope_inv_freq(d).cuda()
    table = L.rope_table(row, col, inv, d)
    wp = pack_w_in(wqkv, w1)
    qp = wp.shape[0] - 2 * Hf
    qkv, act = L.qkv_swiglu(h, wp, D, d, Hf, qp, nq, nk, table)
    # oracle with bf16 tensors = the reference's eager rounding points
    hc = h.cpu()
    r_qkv = F.linear(hc.float(), wqkv.cpu().float()).to(torch.bfloat16).reshape(1, M, 3, heads, d)
    q, k, v = r_qkv.unbind(2)
    q = ae_oracle.rms_norm(q, nq.cpu())
    k = ae_oracle.rms_norm(k, nk.cpu())
    cos, sin = ae_oracle.rope_cos_sin(row.cpu()[None], col.cpu()[None], d)
    q, k = ae_oracle.apply_rope(q, cos, sin), ae_oracle.apply_rope(k, cos, sin)
    got = qkv.cpu().reshape(1, M, 3, heads, d)
    report(f"qkv.q M={M} D={D}", got[:, :, 0], q.float(), max_abs=1.3e-1, rel_fro=8e-3)
    report(f"qkv.k M={M} D={D}", got[:, :, 1], k.float(), max_abs=1.3e-1, rel_fro=8e-3)
    report(f"qkv.v M={M} D={D}", got[:, :, 2], v.float(), max_abs=6e-2, rel_fro=4e-3)
    u = F.linear(hc.float(), w1.cpu().float()).to(torch.bfloat16)
    val, gate = u.chunk(2, dim=-1)
    r_act = F.silu(gate) * val
    report(f"swiglu act M={M} D={D}", act, r_act.float(), max_abs=1.3e-1, rel_fro=8e-3)


@pytest.mark.parametrize("M,D,Hf", [(256, 128, 336), (2048, 1024, 2736), (333, 256, 688), (1024, 3072, 8208)])
def test_proj_residual(L, M, D, Hf):
    a = bf16_randn(M, D + Hf, seed=30)
    w = bf16_randn(D, D + Hf, seed=31, scale=1 / math.sqrt(D + Hf))
    gamma = (torch.rand(D, generator=torch.Generator().manual_seed(32)) + 0.5).to(torch.bfloat16).cuda()
    x = bf16_randn(M, D, seed=33)
    x0 = x.clone()
    L.proj_residual(a, w, gamma, x)
    acc = F.linear(a.cpu().float(), w.cpu().float())
    ref = (x0.cpu() + (acc.to(torch.bfloat16) * gamma.cpu()))
    report(f"proj_residual M={M} D={D}", x, ref.float(), max_abs=1.3e-1, rel_fro=4e-3)   # <= 2 bf16 ulp at |x| < 16


@pytest.mark.parametrize("M,D,Hf", [(2048, 1024, 2736), (1000, 1024, 2736), (1024, 1024, 2736), (264, 1024, 2736), (768, 3072, 8208)])
def test_proj_residual_split_k(L, M, D, Hf):
    """Small batches run the residual GEMM split-K over the two CTA pairs of a 4-CTA cluster (K-halves summed in fp32 through
    distributed shared memory).  Same answer as the un-split kernel up to the fp32 summation order: equal to the fp32
    reference within the usual gate, and the two kernels differ from each other by at most one bf16 ulp on a few elements."""
    a = bf16_randn(M, D + Hf, seed=40)
    w = bf16_randn(D, D + Hf, seed=41, scale=1 / math.sqrt(D + Hf))
    gamma = (torch.rand(D, generator=torch.Generator().manual_seed(42)) + 0.5).to(torch.bfloat16).cuda()
    x0 = bf16_randn(M, D, seed=43)
    outs = {}
    try:
        for flag in (1, 0):
            L.set_flag("gemm_splitk", flag)
            x = x0.clone()
            L.proj_residual(a, w, gamma, x)
            outs[flag] = x
    finally:
        L.set_flag("gemm_splitk", 1)
    acc = F.linear(a.cpu().float(), w.cpu().float())
    ref = (x0.cpu() + (acc.to(torch.bfloat16) * gamma.cpu()))
    report(f"proj_residual split-K M={M} D={D}", outs[1], ref.float(), max_abs=1.3e-1, rel_fro=4e-3)
    report(f"proj_residual un-split M={M} D={D}", outs[0], ref.float(), max_abs=1.3e-1, rel_fro=4e-3)
    diff = (outs[1].float() - outs[0].float()).abs()
    frac = float((diff > 0).float().mean())
    assert frac < 0.05, frac                                  # only accumulators that sat on a bf16 rounding boundary move
    assert float(diff.max()) <= 2 * float(outs[0].float().abs().max()) * 2 ** -8


def test_bad_args_raise(L):
    a = bf16_randn(128, 100, seed=1)   # K not a multiple of 8
    w = bf16_randn(64, 100, seed=2)
    with pytest.raises(ValueError):
        L.linear(a, w)
    with pytest.raises(RuntimeError):
        L.linear(a.cpu(), w.cpu())


@pytest.mark.parametrize("K,M,N", [(1000, 256, 512), (4096, 3072, 1024), (333, 776, 136), (64, 128, 64), (8192, 1024, 2736), (2048, 4648, 1000)])
def test_linear_tn_transposed_operands(K, M, N):
    """out = At^T Bt with both operands given transposed (the training step's weight gradient dW = dY^T X): the tensor core
    reads the [k, column] tiles MN-major, no transpose pass.  Operands are column slices of wider buffers (row pitch > width)."""
    from vitok_b200 import _lib
    at_full = bf16_randn(K, M + 64, seed=81)
    bt_full = bf16_randn(K, N + 24, seed=82)
    at, bt = at_full[:, 32:32 + M], bt_full[:, 8:8 + N]
    out = _lib.linear_tn(at, bt)
    ref = at.float().t() @ bt.float()
    report(f"linear_tn K={K} M={M} N={N}", out, ref, rel_fro=4e-3)


@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (8192, 1024, 3760), (300, 136, 776), (4096, 11280, 3072), (4616, 1032, 1024)])
def test_linear_nn_b_transposed(M, N, K):
    """out = A Bt with B given as [K, N] row-major (the data gradient dX = dY W, W used as stored)."""
    from vitok_b200 import _lib
    a = bf16_randn(M, K + 8, seed=83)[:, :K]
    bt = bf16_randn(K, N + 16, seed=84)[:, 8:8 + N]
    out = _lib.linear_nn(a, bt)
    report(f"linear_nn M={M} N={N} K={K}", out, a.float() @ bt.float(), rel_fro=4e-3)
