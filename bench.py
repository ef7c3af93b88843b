#!/usr/bin/env python
"""bench.py -- encode+decode images/sec of the B200-native ViTok-v2 AE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling auto|weak|strong]
                    [--workload c2|c3|c4|350M-512|350M-2048sw|350M-4096sw|c5|c5-350M]

Contract (see DESIGN.md "Measurement"):
  * one "step" = AE.encode + AE.decode over one batch of synthetic images of the named resolution;
  * workload at N=1 (default c2) = BASELINE.json configs[1]: 350M-f16x64, bf16, 64 x 256x256 images;
  * N > 1 (torchrun, one rank per GPU, no data-path collective): c2 is BASELINE's "batch 64 ... batch-sharded to 2/4/8", i.e.
    STRONG scaling -- the 64-image batch is dealt 64/N images per rank and the step is replayed from a CUDA graph
    (vitok_b200.GraphedAE / GraphedCodec) -- and the line also carries the WEAK-scaling figure (64 images on every rank) under
    "weak"; --scaling weak|strong forces one of them as `value`.  The other workloads keep a fixed per-GPU batch (weak);
  * `value`   : whole-job images/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks;
  * `e2e`     : same metric through the public API with HOST (pinned) buffers: uint8 images in (H2D), preprocess,
                encode, decode, postprocess, uint8 reconstructions out (D2H), all inside the timed region;
  * `roofline`: the dominant kernel (QKV+SwiGLU tcgen05 GEMM): algorithmic FLOPs per launch / its mean launch
                duration from CUDA events recorded around every launch inside a timed pass;
  * `cpu_baseline` / `--impl reference`: the reference's own CPU path on the host cores -- the UNMODIFIED reference from
                baseline/_ref (`kind: "reference"`) when it is installed, else the CPU oracle (oracle/, a restatement of it,
                `kind: "port"`; the only place bench.py touches oracle/) -- on the arm's batch or a bounded sample of it.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))

import torch  # noqa: E402

WORKLOADS = {
    # name: (variant, per-GPU batch, resolution, max_tokens, attn_backend)
    "c2": ("Ld4-Ld24/1x16x64", 64, 256, 256, "flash"),
    "c4": ("Td4-T/1x16x64", 8, 512, 1024, "flash"),
    "350M-512": ("Ld4-Ld24/1x16x64", 16, 512, 1024, "flash"),
    # BASELINE configs[2]: NaFlex mixed-aspect batch (sizes drawn in [128,512]^2, non-multiples of 16 included), masked
    # varlen attention (sdpa backend = the only reference backend that honours patch_mask); resolution 0 = ragged
    "c3": ("Ld4-Ld24/1x16x16", 64, 0, 1024, "sdpa"),
    # very-high-resolution decode of the reference's project page (docs/index.html:1140,1304): 2048 px = 16 384
    # tokens, 4096 px = 65 536 tokens per image; --sw W adds the sliding window AE(sw=W) (ae.py:90,99; attention.py:113-116)
    "350M-2048": ("Ld4-Ld24/1x16x64", 2, 2048, 16384, "flash"),
    "350M-4096": ("Ld4-Ld24/1x16x64", 1, 4096, 65536, "flash"),
    # training-step configs (BASELINE configs[4]): forward + Charbonnier + backward + AdamW, DDP all-reduce when N > 1
    "c5": ("Td4-T/1x32x256", 8, 1024, 1024, "flash"),
    "c5-350M": ("Ld4-Ld24/1x16x64", 16, 256, 256, "flash"),
}
TRAIN_WORKLOADS = ("c5", "c5-350M")
STRONG_GLOBAL_BATCH = {"c2": 64}     # workloads BASELINE.json words as "batch B ... batch-sharded to 2/4/8": strong scaling when N > 1
METRIC = "encode+decode images/sec"
UNIT = "images/s"
CLS_NAMES = ["linear", "rmsnorm", "qkv_swiglu_gemm", "attention", "proj_residual_gemm", "misc"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def c3_sizes(n, seed):
    """Seeded (H, W) list for the c3 workload: uniform in [128, 512], every image fits max_tokens = 1024 at p = 16."""
    import numpy as np
    rng = np.random.RandomState(seed)
    return [(int(rng.randint(128, 513)), int(rng.randint(128, 513))) for _ in range(n)]


def flops_per_image(cfg, N, window=0):
    """per layer 2N D (4D + 3Hf) + 4 N^2 D, plus the four projections.  With a sliding window of radius
    `window` tokens a query sees at most 2 * window + 1 keys (clipped at the sequence ends): sum_i |{j : |i - j| <= window}|."""
    def hf(D):
        return ((int(D * cfg["mlp_factor"]) + 8) // 16) * 16
    pairs = float(N) * N
    if window and window > 0 and window < N:
        w = int(window)
        pairs = float(N) * (2 * w + 1) - float(w) * (w + 1)      # full bands minus the two clipped triangles
    tot = 0.0
    for D, L in ((cfg["encoder_width"], cfg["encoder_depth"]), (cfg["decoder_width"], cfg["decoder_depth"])):
        tot += L * (2.0 * N * D * (4 * D + 3 * hf(D)) + 4.0 * pairs * D)
    P, C, De, Dd = cfg["pixels_per_token"], cfg["channels_per_token"], cfg["encoder_width"], cfg["decoder_width"]
    tot += 2.0 * N * (P * De + De * C + C * Dd + Dd * P)
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if os.environ.get("VTK_BENCH_NO_SAMPLER") == "1":     # debugging aid: no nvidia-smi polling during the timed region
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def n_samples(self) -> int:
        try:
            return sum(1 for r in open(self.f.name) if r.strip())
        except Exception:  # noqa: BLE001
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        if sm:
            load = sorted(sm)[len(sm) // 2:]   # samples under load = upper half
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend, rank=rank, world_size=world,
                                **({"device_id": torch.device("cuda", local)} if backend == "nccl" else {}))
    return rank, world, local


def max_over_ranks(value: float, world: int, device) -> float:
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's PyTorch CPU path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(variant: str, n_images: int, res: int, steps: int, warmup: int, backend: str = "sdpa"):
    from oracle import ae_oracle, pp_oracle
    from oracle.weights import make_state_dict, synth_images
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ae_oracle.decode_variant(variant)
    sd = make_state_dict(cfg, seed=0)
    if res > 0:
        sizes, T = [(res, res)] * n_images, (res // cfg["spatial_stride"]) ** 2
    else:
        sizes, T = c3_sizes(n_images, 1234), 1024
    b = pp_oracle.collate([pp_oracle.patchify(i, cfg["spatial_stride"], T) for i in synth_images(sizes, seed=1234)])
    batch = {k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            enc = ae_oracle.encode(sd, batch, cfg["encoder_heads"], attn_backend=backend)
            ae_oracle.decode(sd, enc, cfg["decoder_heads"], attn_backend=backend)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return n_images * len(times) / total, cores, total / len(times)


def cpu_reference_rate(variant: str, n_images: int, res: int, steps: int, warmup: int):
    """The UNMODIFIED reference (vitok, installed into baseline/_ref -- git-ignored, shipped to the GPU box by gpurun) on the host
    cores: AE(**decode_variant(variant)) in fp32, random init, encode -> decode of the same synthetic batch the port arm uses, through
    the reference's own public API.  attn_backend="sdpa": flash-attn is CUDA-only; with every token valid the two backends compute the
    same attention (attention.py:109-127).  Returns None when baseline/_ref is absent (the caller falls back to the oracle port)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "vitok")):
        return None
    import types
    import numpy as np
    from oracle import pp_oracle
    from oracle.weights import synth_images
    if "webdataset" not in sys.modules:          # the reference's only missing hard import (vitok/data.py), unused on this path
        sys.modules["webdataset"] = types.ModuleType("webdataset")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from vitok.models.ae import AE as RefAE, decode_variant as ref_decode_variant
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ref_decode_variant(variant)
    torch.manual_seed(0)
    model = RefAE(**cfg, attn_backend="sdpa").eval()
    if res > 0:
        sizes, T = [(res, res)] * n_images, (res // cfg["spatial_stride"]) ** 2
    else:
        sizes, T = c3_sizes(n_images, 1234), 1024
    b = pp_oracle.collate([pp_oracle.patchify(i, cfg["spatial_stride"], T) for i in synth_images(sizes, seed=1234)])
    batch = {k: torch.from_numpy(np.asarray(v)) for k, v in b.items()}
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            model.decode(model.encode(batch))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return n_images * len(times) / total, cores, total / len(times)


def _res_name(res):
    return f"{res}px" if res > 0 else "128-512px mixed aspect (NaFlex, max_tokens 1024)"


def resolve_scaling(args, world):
    """('strong' | 'weak', per-GPU batch).  Strong: a fixed global batch dealt to the ranks (BASELINE's "batch-sharded")."""
    variant, B, res, T, backend = WORKLOADS[args.workload]
    if args.batch:
        return ("weak", args.batch) if args.scaling != "strong" else ("strong", max(1, args.batch // world))
    gb = STRONG_GLOBAL_BATCH.get(args.workload)
    mode = args.scaling
    if mode == "auto":
        mode = "strong" if (gb and world > 1) else "weak"
    if mode == "strong":
        gb = gb or B
        if gb % world:
            raise SystemExit(f"bench.py: global batch {gb} is not divisible by {world} GPUs")
        return "strong", gb // world
    return "weak", B


def make_config(args, world, scaling, B):
    """The `config` object of the JSON line -- identical for our arm and the reference arm (same workload, same batch)."""
    variant, _, res, T, backend = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: {variant} encode+decode @{_res_name(res)}, "
                        + (f"global batch {B * world} sharded {B}/GPU" if scaling == "strong" else f"batch {B}/GPU"),
            "variant": variant, "resolution": res, "tokens_per_image": T, "batch_per_gpu": B, "global_batch": B * world,
            "attn_backend": backend, "sliding_window": args.sw or None, "weights": "random init (seed 0)"}


def run_reference(args, rank, world):
    """The reference's CPU path (the oracle port of vitok/models/ae.py, fp32, all host threads) on this arm's config.  A step
    is the config's batch when one step fits in ~10 s of CPU work (c2: the whole 64-image batch dealt as in our arm), else a
    bounded sample of it, stated in cpu_baseline.sample and config.reference_sample."""
    if rank != 0:
        return
    variant, _, res, T, backend = WORKLOADS[args.workload]
    scaling, B = resolve_scaling(args, world)
    est_gflop_img = {256: 189.0, 512: 846.0}.get(res, 1e9) if variant.startswith("L") else 1e9
    full = B * est_gflop_img <= 64 * 189.0 * 1.01            # one step <= the c2 batch (~7 s on 16 cores)
    n_img = B if full else (4 if (0 < res <= 256) else 1)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if not full:            # bounded: the whole run must end within a few minutes; the line prints what was actually run
        steps, warmup = min(steps, 3 if est_gflop_img > 1e6 else 20), min(warmup, 1)
    kind, got = "reference", None
    try:
        got = cpu_reference_rate(variant, n_img, res, steps, warmup)
    except Exception as ex:                                   # a broken install must not cost the arm: say so and use the port
        sys.stderr.write(f"bench.py: baseline/_ref could not be run ({type(ex).__name__}: {ex}); using the oracle port\n")
    if got is None:
        kind, got = "port", cpu_oracle_rate(variant, n_img, res, steps, warmup)
    rate, cores, sec = got
    what = ("the unmodified reference from baseline/_ref: vitok.models.ae.AE, sdpa backend" if kind == "reference"
            else "oracle port of vitok/models/ae.py")
    sample = (f"{n_img} x {_res_name(res)} images per step ({'the whole per-rank batch' if full else 'a bounded sample of the batch'}), "
              f"fp32, torch CPU ops on {cores} threads ({what}), {steps} timed steps after {warmup} warm-up")
    cfg = make_config(args, world, scaling, B)
    if not full:
        cfg["reference_sample"] = f"{n_img} of {B} images per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Inputs:
    """Synthetic batch of one rank: float images, the patch dict resident in HBM and the pinned uint8 host copy."""

    def __init__(self, vb, cfg, wl, B, rank, dev):
        variant, _, res, T, backend = wl
        g = torch.Generator().manual_seed(1234 + rank)
        self.patch = patch = cfg["spatial_stride"]
        self.ragged = res == 0
        self.B, self.T, self.res = B, T, res
        if self.ragged:
            self.sizes = c3_sizes(B, 1234 + rank)
            img_list = [torch.rand(3, h, w, generator=g) * 2 - 1 for h, w in self.sizes]
            self.pd = vb.patchify_batch([i.to(dev) for i in img_list], patch, T, out_dtype=torch.bfloat16, device=dev)
            self.n_valid = [-(-h // patch) * -(-w // patch) for h, w in self.sizes]
            u8_list = [((i.permute(1, 2, 0) + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).contiguous() for i in img_list]
            self.host_u8, self.offs, self.szs = vb.pack_images(u8_list, pin=True)
            self.canvas = 512
        else:
            imgs = torch.rand(B, 3, res, res, generator=g) * 2 - 1
            self.pd = vb.patchify_batch(imgs.to(dev), patch, T, out_dtype=torch.bfloat16, device=dev)
            self.n_valid = [T] * B
            self.host_u8 = ((imgs.permute(0, 2, 3, 1) + 1) * 127.5).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
            self.canvas = res
        self.host_out = [torch.empty(B, 3, self.canvas, self.canvas, dtype=torch.uint8).pin_memory() for _ in range(2)]


def time_resident(vb, model, inp, steps, warmup, world, dev, use_graph, sampler_rank0):
    """Device-resident timing: encode + decode over the HBM-resident patch dict, CUDA events on the launching stream,
    barrier + synchronize on both sides, max over ranks.  Returns (ms total, launches per step, clocks)."""
    def eager():
        with torch.no_grad():
            e = model.encode(inp.pd)
            n = model.last_launch_count
            o = model.decode(e)
            return o, n + model.last_launch_count

    _, per_step = eager()
    join = lambda: None                     # noqa: E731
    if use_graph:
        # Small per-rank batches: every kernel of the chain is about one wave and leaves SMs idle at its head and tail, so TWO steps are
        # kept in flight -- two graphs of the same model with workspaces of their own, replayed alternately on two streams (each step is
        # still one encode + decode of one batch; VTK_BENCH_DUAL_STREAM=0: one graph, one stream).
        n_g = 2 if os.environ.get("VTK_BENCH_DUAL_STREAM", "1") != "0" else 1
        graphs = [vb.GraphedAE(model, inp.pd, private_workspace=n_g > 1) for _ in range(n_g)]
        main = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_g)] if n_g > 1 else [main]
        turn = [0]

        def run():
            i = turn[0] % n_g
            turn[0] += 1
            if n_g > 1:
                streams[i].wait_stream(main)                 # (orders the replay after the timing event recorded on `main`)
            with torch.cuda.stream(streams[i]):
                graphs[i](graphs[i].static_in)               # static inputs: no copies, one graph launch

        def join():
            for st in streams:
                if st is not main:
                    main.wait_stream(st)
    else:
        run = lambda: eager()[0]            # noqa: E731
    for _ in range(max(warmup, 3)):
        run()
    join()
    torch.cuda.synchronize()
    barrier(world)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index) if sampler_rank0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        run()
    join()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = ev0.elapsed_time(ev1)
    if sampler is not None:
        # nvidia-smi needs a few hundred ms to start: if the timed region was shorter than that, keep the same load
        # running (untimed) until the sampler has seen it, so that `clocks` always describes the GPU under this workload
        t_end = time.perf_counter() + 3.0
        while sampler.n_samples() < 4 and time.perf_counter() < t_end:
            for _ in range(4):
                run()
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    return max_over_ranks(ms, world, dev), per_step, clocks


def time_e2e(vb, model, inp, steps, world, dev, use_graph):
    """The serving loop a user of the reference writes (README.md:62-65): decoded uint8 images on the host -> preprocess
    (to_tensor|normalize|patchify) -> encode -> decode -> postprocess (unpatchify, 0_255) -> uint8 images on the host.
    Per step: H2D of the uint8 HWC batch, D2H of the uint8 CHW reconstructions, both inside the timed region, double-buffered
    on copy-in / compute / copy-out streams so that step i's copies overlap step i+-1's kernels.  With use_graph the device
    work of a step is one replay of vitok_b200.GraphedCodec (two instances = the two buffers)."""
    B, patch, T, canvas = inp.B, inp.patch, inp.T, inp.canvas
    host_u8 = inp.host_u8
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    s_main = torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    diag = {"h2d": [], "d2h": [], "cpu": []}
    dual = use_graph and os.environ.get("VTK_BENCH_DUAL_STREAM", "1") != "0"    # (see time_resident: two steps in flight)
    s_comp = [torch.cuda.Stream(device=dev) for _ in range(2)] if dual else [s_main, s_main]
    if use_graph:
        first = host_u8.to(dev)
        codecs = [vb.GraphedCodec(model, (first, inp.offs, inp.szs) if inp.ragged else first, patch, T,
                                  max_grid_size=canvas // patch, output_format="0_255", private_workspace=dual) for _ in range(2)]
        dev_in = [c.static_in for c in codecs]
    else:
        dev_in = [torch.empty_like(host_u8, device=dev) for _ in range(2)]
    for e in ev_free + ev_copied:
        e.record(s_main)
    keep = [None, None]

    def step(i, timed=False):
        b = i & 1
        c0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])                 # the patchify of step i-2 has consumed this input buffer
            if timed:
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record(s_in)
            dev_in[b].copy_(host_u8, non_blocking=True)
            if timed:
                h1.record(s_in)
                diag["h2d"].append((h0, h1))
            ev_in[b].record(s_in)
        sc = s_comp[b]                                  # the compute stream of this step (dual: buffer b has its own)
        sc.wait_event(ev_in[b])
        if use_graph:
            sc.wait_event(ev_copied[b])                 # the D2H copy of step i-2 has read this codec's output buffer
            with torch.cuda.stream(sc):
                img = codecs[b]()
            ev_free[b].record(sc)
        else:
            if inp.ragged:
                d = vb.patchify_packed(dev_in[b], inp.offs, inp.szs, patch, T, out_dtype=torch.bfloat16)
            else:
                d = vb.patchify_batch(dev_in[b], patch, T, out_dtype=torch.bfloat16, device=dev)
            ev_free[b].record(s_main)
            with torch.no_grad():
                o = model.decode(model.encode(d))
            img = vb.unpatchify(o, patch, max_grid_size=canvas // patch, output_format="0_255")
            # the D2H copy of step i-2 (it read keep[b]) has finished before that tensor's memory can be reused on s_main
            s_main.wait_event(ev_copied[b])
            keep[b] = img
        ev_done[b].record(sc)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[b])
            if timed:
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                d0.record(s_out)
            inp.host_out[b].copy_(img, non_blocking=True)
            if timed:
                d1.record(s_out)
                diag["d2h"].append((d0, d1))
            ev_copied[b].record(s_out)
        if timed:
            diag["cpu"].append((time.perf_counter() - c0) * 1e3)

    def drain():
        for st in s_comp:
            if st is not s_main:
                s_main.wait_stream(st)
        s_main.wait_stream(s_out)
        s_main.wait_stream(s_in)

    for i in range(4):
        step(i)
    drain()
    torch.cuda.synchronize()
    barrier(world)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        step(i, timed=True)
    drain()
    t1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(t0.elapsed_time(t1), world, dev)
    return {"value": world * B * steps / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": host_u8.numel(),
            "d2h_bytes_per_step": inp.host_out[0].numel(), "ms_per_step": ms / steps,
            "h2d_copy_ms": statistics.mean(a.elapsed_time(b) for a, b in diag["h2d"]),
            "d2h_copy_ms": statistics.mean(a.elapsed_time(b) for a, b in diag["d2h"]),
            "cpu_enqueue_ms_per_step": statistics.mean(diag["cpu"]), "cuda_graph": bool(use_graph),
            "launches_per_step": codecs[0].launches if use_graph else None}


def kernel_breakdown(lib, model, cfg, inp, steps, args):
    """Per-kernel-class CUDA-event timing inside a timed (eager) pass + the roofline object of the dominant kernel."""
    import ctypes
    h = model._handle
    lib.vtk_ae_set_timing(h, 1)
    ms_cls, cnt_cls = (ctypes.c_float * 6)(), (ctypes.c_int * 6)()
    tot, cnt = [0.0] * 6, [0] * 6
    prof_steps = min(steps, 10)
    for _ in range(prof_steps):
        with torch.no_grad():
            e = model.encode(inp.pd)
            lib.vtk_ae_collect_timing(h, ms_cls, cnt_cls)
            for i in range(6):
                tot[i] += ms_cls[i]; cnt[i] += cnt_cls[i]
            model.decode(e)
            lib.vtk_ae_collect_timing(h, ms_cls, cnt_cls)
            for i in range(6):
                tot[i] += ms_cls[i]; cnt[i] += cnt_cls[i]
    lib.vtk_ae_set_timing(h, 0)
    breakdown = {CLS_NAMES[i]: {"ms_per_step": tot[i] / prof_steps, "launches_per_step": cnt[i] // prof_steps} for i in range(6)}
    # dominant kernel: QKV+SwiGLU GEMM; FLOPs per launch averaged over the encoder + decoder launches
    M = sum(inp.n_valid)    # algorithmic rows: valid tokens only (padded tokens are not work)
    fl = 0.0
    for D, L in ((cfg["encoder_width"], cfg["encoder_depth"]), (cfg["decoder_width"], cfg["decoder_depth"])):
        hf = ((int(D * cfg["mlp_factor"]) + 8) // 16) * 16
        fl += L * 2.0 * M * D * (3 * D + 2 * hf)
    n_l = cfg["encoder_depth"] + cfg["decoder_depth"]
    avg_ms = tot[2] / max(cnt[2], 1)
    pk = peaks()
    ach = (fl / n_l) / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    # dram bytes per launch of this kernel from the newest committed `ncu --set full` capture of this build's kernel at this
    # workload's full batch (profiles/rNN_traffic.json, written by tools/ncu_summary.py); null for any other shape
    traffic, tsrc = None, None
    for name in sorted((f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json")), reverse=True):
        ent = json.load(open(os.path.join(ROOT, "profiles", name))).get(args.workload, {})
        if ent.get("qkv_swiglu_gemm_dram_bytes_per_launch") and ent.get("batch", WORKLOADS[args.workload][1]) == inp.B:
            traffic, tsrc = ent["qkv_swiglu_gemm_dram_bytes_per_launch"], name
            break
    # --quantize: no FP8 peak was measured on this pool; the dense e4m3 rate of the tensor core is twice the bf16 one, so the
    # denominator is 2 x the measured sustained bf16 figure (stated in peak_source); traffic is the bf16 capture's and is dropped
    peak = pk["bf16_sustained"] * (2.0 if args.quantize else 1.0)
    roof = {"kernel": "gemm2_kernel<EPI_QKV_SWIGLU> (cta_group::2, 256x256 pair tile" + (", e4m3 operands)" if args.quantize else ")"),
            "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": None if args.quantize else traffic, "traffic_source": tsrc,
            "peak_source": pk["src"] + (" sustained bf16 x 2 (dense e4m3 : bf16 rate)" if args.quantize else " sustained bf16"),
            "avg_launch_ms": avg_ms, "flops_per_launch": fl / n_l}
    return roof, breakdown


def run_ours(args, rank, world, local):
    import vitok_b200 as vb
    from vitok_b200 import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    variant, B_weak, res, T, backend = wl
    scaling, B = resolve_scaling(args, world)
    cfg = vb.decode_variant(variant)
    torch.manual_seed(0)
    model = vb.AE(**cfg, attn_backend=backend, sw=(args.sw or None)).eval().to(device=dev, dtype=torch.bfloat16)
    if args.quantize:
        model.quantize()
    lib = _lib.load()
    # CUDA-graph replay of the step: at small per-GPU batches the ~100 launches of a step take the host longer than the GPU
    # needs to run them (B = 8: 2 ms of GPU work, 5 ms of launches).  auto = whenever the per-GPU batch is below 32 images.
    use_graph = args.graph == "on" or (args.graph == "auto" and B * T < 32 * 256)
    inp = Inputs(vb, cfg, wl, B, rank, dev)
    ms, launches_per_step, clocks = time_resident(vb, model, inp, args.steps, args.warmup, world, dev, use_graph, rank == 0)
    value = world * B * args.steps / (ms / 1e3)
    e2e = time_e2e(vb, model, inp, args.steps, world, dev, use_graph)
    roof, breakdown = kernel_breakdown(lib, model, cfg, inp, args.steps, args)

    # the other scaling mode of this workload, measured in the same run (c2, N > 1: `value` is the strong-scaling figure -- 64
    # images dealt to the ranks -- and "weak" is 64 images on every rank; --scaling weak swaps them)
    other = None
    if world > 1 and args.workload in STRONG_GLOBAL_BATCH and not args.batch and args.both:
        o_scaling = "weak" if scaling == "strong" else "strong"
        oB = B_weak if o_scaling == "weak" else STRONG_GLOBAL_BATCH[args.workload] // world
        o_graph = args.graph == "on" or (args.graph == "auto" and oB * T < 32 * 256)
        o_inp = Inputs(vb, cfg, wl, oB, rank, dev)
        o_ms, _, _ = time_resident(vb, model, o_inp, args.steps, args.warmup, world, dev, o_graph, False)
        o_e2e = time_e2e(vb, model, o_inp, args.steps, world, dev, o_graph)
        other = {"scaling": o_scaling, "value": world * oB * args.steps / (o_ms / 1e3), "unit": UNIT, "ms_per_step": o_ms / args.steps,
                 "batch_per_gpu": oB, "global_batch": oB * world, "cuda_graph": bool(o_graph),
                 "e2e": {k: o_e2e[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}}
    if rank != 0:
        return
    gf = sum(flops_per_image(cfg, n, args.sw) for n in inp.n_valid) / B / 1e9     # per image, valid tokens only
    pk = peaks()
    try:
        n_img = 4 if 0 < res <= 256 else 1
        c_res = res if res <= 512 else 512                                   # bounded sample: the CPU leg stops at 512 px
        kind, got = "reference", None
        try:                                                                 # the unmodified reference (baseline/_ref) when it is installed
            got = cpu_reference_rate(variant, n_img, c_res, 2, 1)
        except Exception as ex:
            sys.stderr.write(f"bench.py: baseline/_ref could not be run ({type(ex).__name__}: {ex}); using the oracle port\n")
        if got is None:
            kind, got = "port", cpu_oracle_rate(variant, n_img, c_res, 2, 1)
        rate, cores, sec = got
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_img} x {_res_name(c_res)} images, fp32, 2 timed iterations after 1 warm-up ({sec:.2f} s each)"}
    except Exception as ex:  # noqa: BLE001
        cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    config = make_config(args, world, scaling, B)
    config.update({"valid_tokens_per_gpu": sum(inp.n_valid), "token_packing": bool(inp.ragged), "cuda_graph": bool(use_graph),
                   "l2": "no flush: per-step working set (weights + activations) exceeds the 126 MB L2", "gflop_per_image": gf})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "fp8 e4m3 block GEMMs (AE.quantize), bf16 elsewhere" if args.quantize else "bf16",
        "data": "synthetic", "config": config, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "model_tflops": value * gf / 1e3 / world,
        "model_frac_of_peak": value * gf / 1e3 / world / pk["bf16_sustained"],
        "roofline": roof, "kernel_breakdown": breakdown, "cpu_baseline": cpu, "clocks": clocks,
    }
    if other is not None:
        line[other["scaling"]] = other
    print(json.dumps(line), flush=True)


def run_train(args, rank, world, local):
    """Training step (scripts/train_vae.py:304-320,371-372): model(batch) -> Charbonnier -> backward -> AdamW, wrapped in
    DDP (train_vae.py:172) when world > 1 so the gradient all-reduce runs over NCCL."""
    import vitok_b200 as vb
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    variant, B, res, T, backend = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    cfg = vb.decode_variant(variant)
    torch.manual_seed(0)
    with torch.device(dev):
        model = vb.AE(**cfg, attn_backend=backend)
    model = model.to(torch.bfloat16).train()
    net = model
    sync_mode = "none"
    if world > 1:
        if os.environ.get("VTK_TRAIN_DDP", "0") == "1":     # torch DDP (what scripts/train_vae.py:172 uses): all-reduce after backward
            from torch.nn.parallel import DistributedDataParallel as DDP
            net = DDP(model, device_ids=[local], static_graph=True, gradient_as_bucket_view=True)
            sync_mode = "torch DDP"
        else:                                                # all-reduce issued per block from inside the backward loop
            vb.enable_grad_sync(model)
            sync_mode = "overlapped NCCL all-reduce (vb.enable_grad_sync)"
    decay = [p for n, p in model.named_parameters() if not (p.ndim <= 1 or "bias" in n or "norm" in n or "embedding" in n)]
    no_decay = [p for n, p in model.named_parameters() if (p.ndim <= 1 or "bias" in n or "norm" in n or "embedding" in n)]
    opt = vb.FusedAdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-4, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(1234 + rank)
    imgs = torch.rand(B, 3, res, res, generator=g) * 2 - 1
    pd = vb.patchify_batch(imgs.to(dev), cfg["spatial_stride"], T, out_dtype=torch.bfloat16, device=dev)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        out = net(pd)
        loss = vb.charbonnier_loss(out["patches"], pd["patches"], pd["patch_mask"], eps=1e-3)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 1)):
        losses.append(float(step().detach()))
    torch.cuda.synchronize()
    barrier(world)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0.record()
    for _ in range(args.steps):
        last = step()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    losses.append(float(last.detach()))
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ev0.elapsed_time(ev1), world, dev)
    # exposed (non-overlapped) all-reduce time: the same steps once more with the gradient exchange switched off (ranks drift apart from
    # here on -- these are the last steps of the run); exposed = synchronised step - local-only step, both max over ranks
    exposed = None
    if world > 1:
        import contextlib
        grp = getattr(model, "_grad_sync_group", None)
        if grp is not None:
            del model._grad_sync_group
        ctx = net.no_sync() if hasattr(net, "no_sync") else contextlib.nullcontext()
        n_local = max(2, min(args.steps, 5))
        with ctx:
            step()
            torch.cuda.synchronize()
            barrier(world)
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for _ in range(n_local):
                step()
            l1.record()
            torch.cuda.synchronize()
        barrier(world)
        local_ms = max_over_ranks(l0.elapsed_time(l1), world, dev) / n_local
        exposed = {"local_only_ms_per_step": local_ms, "allreduce_exposed_ms_per_step": ms / args.steps - local_ms,
                   "gradient_bytes_per_step": sum(p.numel() for p in model.parameters()) * 2}
    if rank != 0:
        return
    N = T
    gf = 3.0 * flops_per_image(cfg, N) / 1e9
    value = world * B * args.steps / (ms / 1e3)
    pk = peaks()
    line = {
        "metric": "training step images/sec", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {variant} training step @{res}px, batch {B}/GPU, Charbonnier + AdamW"
                               + (f", gradient sync: {sync_mode}" if world > 1 else ""),
                   "variant": variant, "resolution": res, "tokens_per_image": N, "batch_per_gpu": B, "global_batch": B * world,
                   "attn_backend": backend, "gflop_per_image_fwd_bwd": gf},
        "model_tflops": value * gf / 1e3 / world, "model_frac_of_peak": value * gf / 1e3 / world / pk["bf16_sustained"],
        "loss_first_last": [losses[0], losses[-1]], "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "clocks": clocks,
        "optimizer": "FusedAdamW: fp32 master weights + fp32 moments (scripts/train_vae.py:200-208 precision), bf16 model weights / gradients",
        "grad_sync": exposed,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="auto: strong for c2 at N > 1 (BASELINE: batch 64 sharded to 2/4/8), weak otherwise")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="replay the step from a CUDA graph (auto: small per-GPU batches)")
    ap.add_argument("--no-both", dest="both", action="store_false", help="c2 at N > 1: skip the second (other scaling mode) measurement")
    ap.add_argument("--sw", type=int, default=0, help="sliding-window radius in tokens, AE(sw=W) (flash backend)")
    ap.add_argument("--quantize", action="store_true",
                    help="AE.quantize(): FP8 (e4m3) block GEMMs -- a separate, reduced-precision line; the headline stays bf16")
    args = ap.parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    try:
        if args.workload in TRAIN_WORKLOADS:
            run_train(args, rank, world, local)
        else:
            run_ours(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
