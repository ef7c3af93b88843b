"""ncu driver: the attention kernel alone on random q/k/v at the c2 and c4 shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch  # noqa: E402
from vitok_b200 import _lib  # noqa: E402

shapes = [(64, 256, 16, 64), (16, 1024, 16, 64), (8, 1024, 24, 128)]
for (B, N, h, d) in shapes:
    qkv = (torch.randn(B * N, 3 * h * d, generator=torch.Generator().manual_seed(0)) * 1.0).to(torch.bfloat16).cuda()
    for _ in range(3):
        out = _lib.attention(qkv, B, N, h, d, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = _lib.attention(qkv, B, N, h, d, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * N * N * h * d * B
    print(f"attn B={B} N={N} h={h} d={d}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
