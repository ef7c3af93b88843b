"""GPU parity: tcgen05 varlen attention vs the CPU oracle."""
import pytest
import torch

from _util import bf16_randn, report
from oracle import ae_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from vitok_b200 import _lib
    _lib.load()
    return _lib


def _ref(qkv, B, N, heads, d, mask):
    t = qkv.cpu().reshape(B, N, 3, heads, d)
    q, k, v = t.unbind(2)
    return ae_oracle.attention_core(q, k, v, mask.cpu() if mask is not None else None).reshape(B * N, heads * d)


@pytest.mark.parametrize("B,N,heads,d", [(2, 128, 2, 64), (2, 256, 4, 64), (1, 1024, 2, 64), (2, 256, 2, 128),
                                         (1, 1024, 3, 128), (3, 64, 2, 64), (2, 200, 2, 64), (1, 384, 1, 128)])
def test_attention_dense(L, B, N, heads, d):
    qkv = bf16_randn(B * N, 3 * heads * d, seed=40)
    out = L.attention(qkv, B, N, heads, d, None)
    report(f"attn dense B={B} N={N} h={heads} d={d}", out, _ref(qkv, B, N, heads, d, None).float(), max_abs=3e-2, rel_fro=1e-2)


@pytest.mark.parametrize("N,heads,d,lens", [(256, 2, 64, [256, 100, 1, 129]), (1024, 2, 64, [64, 1024, 600, 130]),
                                            (512, 2, 128, [512, 77, 256, 300]), (64, 2, 64, [24, 64, 32])])
def test_attention_prefix_mask(L, N, heads, d, lens):
    B = len(lens)
    qkv = bf16_randn(B * N, 3 * heads * d, seed=41)
    mask = torch.zeros(B, N, dtype=torch.bool)
    for b, n in enumerate(lens):
        mask[b, :n] = True
    out = L.attention(qkv, B, N, heads, d, mask.cuda())
    ref = _ref(qkv, B, N, heads, d, mask)
    report(f"attn masked N={N} d={d}", out, ref.float(), max_abs=3e-2, rel_fro=1e-2)
    # padded query rows are defined as 0 here
    assert out.cpu().reshape(B, N, -1)[~mask].abs().max().item() == 0.0


def test_attention_general_mask(L):
    B, N, heads, d = 2, 256, 2, 64
    qkv = bf16_randn(B * N, 3 * heads * d, seed=42)
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(B, N, generator=g) > 0.4
    out = L.attention(qkv, B, N, heads, d, mask.cuda())
    report("attn general mask", out, _ref(qkv, B, N, heads, d, mask).float(), max_abs=3e-2, rel_fro=1e-2)


def test_masked_equals_unpadded(L):
    """Property: a padded image's valid rows == the same image run alone without padding."""
    heads, d, n = 2, 64, 100
    small = bf16_randn(n, 3 * heads * d, seed=43)
    alone = L.attention(small.contiguous(), 1, n, heads, d, None)
    N = 256
    padded = torch.zeros(N, 3 * heads * d, dtype=torch.bfloat16, device="cuda")
    padded[:n] = small
    padded[n:] = bf16_randn(N - n, 3 * heads * d, seed=44)   # garbage in the padding must not matter
    mask = torch.zeros(1, N, dtype=torch.bool)
    mask[0, :n] = True
    out = L.attention(padded, 1, N, heads, d, mask.cuda())
    assert torch.equal(out[:n], alone)
