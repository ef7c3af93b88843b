"""Attention backward alone at a BASELINE training shape (default c5: 8 x 24 heads x N = 1024, d = 128): CUDA-event time of the two
passes (MODE 0 = dK / dV, MODE 1 = dQ) and their tensor-core rate; the driver for `ncu -k regex:attn_bwd`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch  # noqa: E402
from vitok_b200 import _lib as L  # noqa: E402

B, N, heads, d = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 1024, 24, 128)
reps = 5
D = heads * d
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * N, 3 * D, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
do = (torch.randn(B * N, D, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
lse = torch.empty(B * N, heads, dtype=torch.float32, device="cuda")
out = L.attention(qkv, B, N, heads, d, None, window=-1, lse=lse)
delta = torch.empty(B * N, heads, dtype=torch.float32, device="cuda")
L.check(L.load().vtk_attn_delta(out.data_ptr(), D, do.data_ptr(), D, delta.data_ptr(), B * N, heads, d, L.stream_ptr()))
dqkv = torch.empty(B * N, 3 * D, dtype=torch.bfloat16, device="cuda")
b0, g0 = qkv.data_ptr(), dqkv.data_ptr()


def run():
    L.check(L.load().vtk_attention_bwd_bf16(b0, b0 + 2 * D, b0 + 4 * D, 3 * D, do.data_ptr(), D, lse.data_ptr(), delta.data_ptr(),
                                            g0, g0 + 2 * D, g0 + 4 * D, 3 * D, None, B, N, heads, d, 0, -1, L.stream_ptr()))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 7 * 2.0 * B * heads * N * N * d      # 7 MMAs of N x N x d per (image, head): S and dP twice, dV, dK, dQ
print(f"attention backward B={B} N={N} heads={heads} d={d}: {ms * 1e3:.1f} us for both passes, {fl / ms / 1e9:.0f} TFLOP/s issued "
      f"({5 * fl / 7 / ms / 1e9:.0f} algorithmic)")
