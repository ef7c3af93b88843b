"""Per-kernel CUDA time of one training step (bench.py c5 / c5-350M workload) with torch.profiler (CUPTI, no replay)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch  # noqa: E402
import vitok_b200 as vb  # noqa: E402
from bench import WORKLOADS  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c5-350M"
variant, B, res, T, backend = WORKLOADS[wl]
dev = torch.device("cuda", 0)
cfg = vb.decode_variant(variant)
torch.manual_seed(0)
with torch.device(dev):
    model = vb.AE(**cfg, attn_backend=backend)
model = model.to(torch.bfloat16).train()
decay = [p for n, p in model.named_parameters() if not (p.ndim <= 1 or "bias" in n or "norm" in n)]
no_decay = [p for n, p in model.named_parameters() if (p.ndim <= 1 or "bias" in n or "norm" in n)]
opt = vb.FusedAdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-4, betas=(0.9, 0.99))
imgs = torch.rand(B, 3, res, res, generator=torch.Generator().manual_seed(1)) * 2 - 1
pd = vb.patchify_batch(imgs.to(dev), cfg["spatial_stride"], T, out_dtype=torch.bfloat16, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(pd)
    loss = vb.charbonnier_loss(out["patches"], pd["patches"], pd["patch_mask"], eps=1e-3)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"{wl}: step {e0.elapsed_time(e1):.2f} ms")
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
print(f"{wl}: 3 back-to-back steps {e0.elapsed_time(e1) / 3:.2f} ms/step")
if len(sys.argv) > 2 and sys.argv[2] == "noprof":
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
for r in rows[:26]:
    print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}%  n={r.count:5d}  {r.key[:110]}")
print(f"total device time {tot / 1e3:.2f} ms")
