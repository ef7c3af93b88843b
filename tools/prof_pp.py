"""Timing driver for the HBM-bound kernels at the c2 shapes (64 x 256x256): patchify (fp32 CHW / uint8 HWC -> bf16
patches + indices), unpatchify (bf16 patches -> fp32-free uint8 / bf16 canvas), RMSNorm, token pack / unpack.
Prints achieved GB/s against the algorithmic bytes of DESIGN.md section 3.3.  An L2 flush (write of a 256 MB buffer)
separates the timed launches so every number is an HBM number, not an L2 one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch  # noqa: E402
import vitok_b200 as vb  # noqa: E402
from vitok_b200 import _lib  # noqa: E402

dev = "cuda"
B, R, p, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 256, 16, 256     # B = 64: the c2 batch; B = 512: long enough to hide launch + ramp
print(f"--- batch {B} x {R}x{R}")
g = torch.Generator().manual_seed(0)
img_f = (torch.rand(B, 3, R, R, generator=g) * 2 - 1).to(dev)
img_u = (torch.rand(B, R, R, 3, generator=g) * 255).to(torch.uint8).to(dev)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3   # us


def line(name, us, nbytes):
    print(f"{name:46s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / us / 1e3:7.1f} GB/s")


idx = B * T * 25
pd = vb.patchify_batch(img_f, p, T, out_dtype=torch.bfloat16)
line("patchify fp32 CHW -> bf16", timed(lambda: vb.patchify_batch(img_f, p, T, out_dtype=torch.bfloat16)), img_f.numel() * 4 + pd["patches"].numel() * 2 + idx)
line("patchify fp32 CHW -> fp32", timed(lambda: vb.patchify_batch(img_f, p, T, out_dtype=torch.float32)), img_f.numel() * 8 + idx)
line("patchify uint8 HWC -> bf16", timed(lambda: vb.patchify_batch(img_u, p, T, out_dtype=torch.bfloat16)), img_u.numel() + pd["patches"].numel() * 2 + idx)
line("unpatchify bf16 -> uint8 (0_255)", timed(lambda: vb.unpatchify(pd, p, max_grid_size=R // p, output_format="0_255")), pd["patches"].numel() * 2 + B * 3 * R * R + B * T * 17)
line("unpatchify bf16 -> bf16", timed(lambda: vb.unpatchify(pd, p, max_grid_size=R // p)), pd["patches"].numel() * 4 + B * T * 17)
M, D = B * T, 1024
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
w = torch.ones(D, device=dev, dtype=torch.bfloat16)
line("rmsnorm [16384, 1024] bf16", timed(lambda: _lib.rmsnorm(x, w)), 2 * M * D * 2)
mask = torch.rand(B, 1024, generator=g) > 0.5
xx = torch.randn(B, 1024, 768, generator=g).to(torch.bfloat16).to(dev)
plan = _lib.pack_plan(mask.to(dev))
nv = int(mask.sum())
packed = _lib.pack_rows(xx, plan)
mask_d = mask.to(dev)
line("pack_plan [64, 1024] mask (3 launches)", timed(lambda: _lib.pack_plan(mask_d)), B * 1024 * 9)
# (the C entry point into a pre-allocated buffer: `_lib.pack_rows` also zero-fills a fresh output tensor, a second pass over it)
cap = plan["src"].numel()
packed_out = torch.zeros(cap, 768, dtype=torch.bfloat16, device=dev)
line("pack_rows  (P = 768, half the tokens valid)",
     timed(lambda: _lib.check(_lib.load().vtk_pack_rows(_lib.ptr(xx), 768, _lib.ptr(plan["src"]), _lib.ptr(plan["cu"]), B, cap, _lib.ptr(packed_out),
                                                        768, 768, _lib.stream_ptr()))), nv * 768 * 4)
line("unpack_rows (P = 768)", timed(lambda: _lib.unpack_rows(packed, plan, B, 1024)), nv * 768 * 2 + B * 1024 * 768 * 2)
