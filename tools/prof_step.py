"""Tiny driver for ncu: builds the workload of bench.py and runs a few encode+decode steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch  # noqa: E402
import vitok_b200 as vb  # noqa: E402
from bench import WORKLOADS  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant, B, res, T, backend = WORKLOADS[wl]
if len(sys.argv) > 3:          # optional batch override (the per-GPU shards of a strong-scaled job: 8 / 16 / 32)
    B = int(sys.argv[3])
cfg = vb.decode_variant(variant)
torch.manual_seed(0)
model = vb.AE(**cfg, attn_backend=backend).eval().to("cuda", torch.bfloat16)
if os.environ.get("VTK_PROF_QUANTIZE") == "1":      # FP8 block GEMMs (AE.quantize)
    model.quantize()
imgs = (torch.rand(B, 3, res, res, generator=torch.Generator().manual_seed(1234)) * 2 - 1).cuda()
pd = vb.patchify_batch(imgs, cfg["spatial_stride"], T, out_dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(steps):
        out = model.decode(model.encode(pd))
torch.cuda.synchronize()
print("ok", out["patches"].float().abs().mean().item())
