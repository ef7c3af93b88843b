"""Debug driver: the e2e loop of bench.py (c2) repeated, with per-step CPU enqueue time and GPU completion time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
import torch
import vitok_b200 as vb

mode = sys.argv[1] if len(sys.argv) > 1 else "record_stream"
dev = torch.device("cuda", 0)
cfg = vb.decode_variant("Ld4-Ld24/1x16x64")
torch.manual_seed(0)
model = vb.AE(**cfg, attn_backend="flash").eval().to(device=dev, dtype=torch.bfloat16)
B, res, T, patch = 64, 256, 256, 16
imgs = torch.rand(B, 3, res, res) * 2 - 1
host_u8 = ((imgs.permute(0, 2, 3, 1) + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
host_outs = [torch.empty(B, 3, res, res, dtype=torch.uint8).pin_memory() for _ in range(2)]
s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
s_main = torch.cuda.current_stream(dev)
dev_in = [torch.empty_like(host_u8, device=dev) for _ in range(2)]
ev_in = [torch.cuda.Event() for _ in range(2)]
ev_free = [torch.cuda.Event() for _ in range(2)]
ev_done = [torch.cuda.Event() for _ in range(2)]
ev_copied = [torch.cuda.Event() for _ in range(2)]
keep = [None, None]
for e in ev_free + ev_copied:
    e.record(s_main)


def step(i):
    b = i & 1
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev_free[b])
        dev_in[b].copy_(host_u8, non_blocking=True)
        ev_in[b].record(s_in)
    s_main.wait_event(ev_in[b])
    d = vb.patchify_batch(dev_in[b], patch, T, out_dtype=torch.bfloat16, device=dev)
    ev_free[b].record(s_main)
    with torch.no_grad():
        o = model.decode(model.encode(d))
    img = vb.unpatchify(o, patch, max_grid_size=res // patch, output_format="0_255")
    ev_done[b].record(s_main)
    if mode == "keep":
        s_main.wait_event(ev_copied[b])     # the copy that read keep[b] (two steps ago) is done before its memory is reused
        keep[b] = img
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_done[b])
        if mode == "record_stream":
            img.record_stream(s_out)
        host_outs[b].copy_(img, non_blocking=True)
        ev_copied[b].record(s_out)


for i in range(6):
    step(i)
torch.cuda.synchronize()
for rep in range(6):
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    cpu = []
    t0 = time.perf_counter()
    marks[0].record()
    for i in range(20):
        c0 = time.perf_counter()
        step(i)
        cpu.append((time.perf_counter() - c0) * 1e3)
        marks[i + 1].record()
    s_main.wait_stream(s_out); s_main.wait_stream(s_in)
    end = torch.cuda.Event(enable_timing=True); end.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    gpu = [marks[i].elapsed_time(marks[i + 1]) for i in range(20)]
    print(f"[{mode}] rep {rep}: total {marks[0].elapsed_time(end):7.2f} ms (wall {wall:7.2f}) | cpu/step max {max(cpu):5.2f} mean {sum(cpu)/20:5.2f} | "
          f"gpu/step max {max(gpu):6.2f} min {min(gpu):6.2f} | mem reserved {torch.cuda.memory_reserved() >> 20} MB")

if len(sys.argv) > 2 and sys.argv[2] == "cpu":
    import cProfile, pstats
    cpu = []
    for i in range(20):
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        step(i)
        cpu.append((time.perf_counter() - c0) * 1e3)
    torch.cuda.synchronize()
    print(f"pure enqueue cost per step (queue empty): mean {sum(cpu)/20:.2f} ms, min {min(cpu):.2f}, max {max(cpu):.2f}")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        torch.cuda.synchronize()
        step(i)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
