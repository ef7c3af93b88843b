"""CPU tests of the N > 1 host logic with a world_size-2 gloo group (no GPU): how images are dealt to ranks, the
barrier + max-over-ranks timing of bench.py, the reference arm under torchrun-style environments, and the bucket logic of
the overlapped gradient all-reduce (vitok_b200/train.py: _GradSync) on CPU tensors."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fn_name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _run2(fn_name, port):
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, fn_name, ret), nprocs=2, join=True)
    return dict(ret)


# ---------------------------------------------------------------- sharding (pure host logic)
def test_shard_range_and_token_balanced_sharding():
    import numpy as np
    from vitok_b200.parallel import image_cost, shard_by_tokens, shard_range
    for n, w in [(64, 8), (10, 4), (3, 8), (0, 2)]:
        parts = [list(shard_range(n, r, w)) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    rng = np.random.RandomState(0)
    tokens = [int(-(-rng.randint(128, 513) // 16) * -(-rng.randint(128, 513) // 16)) for _ in range(64)]   # the c3 size law
    for w in (2, 4, 8):
        parts = shard_by_tokens(tokens, w)
        assert sorted(sum(parts, [])) == list(range(64))
        loads = [sum(image_cost(tokens[i]) for i in p) for p in parts]
        naive = [sum(image_cost(tokens[i]) for i in shard_range(64, r, w)) for r in range(w)]
        assert max(loads) / (sum(loads) / w) < 1.03                      # LPT: within 3 % of perfect balance
        assert max(loads) <= max(naive) + 1e-6                          # never worse than the contiguous split
        assert parts == shard_by_tokens(tokens, w)                      # deterministic: every rank computes the same deal
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


# ---------------------------------------------------------------- bench.py timing helpers under gloo
def _bench_helpers(rank, world):
    import bench
    bench.barrier(world)
    t = bench.max_over_ranks(10.0 + rank, world, torch.device("cpu"))
    from vitok_b200.parallel import shard_range
    mine = list(shard_range(7, rank, world))
    got = [None] * world
    dist.all_gather_object(got, mine)
    return t, got


def test_bench_max_over_ranks_and_barrier_gloo():
    out = _run2("_bench_helpers", 29611)
    assert out[0][0] == out[1][0] == 11.0                               # max over ranks, same on every rank
    assert sorted(sum(out[0][1], [])) == list(range(7))


# ---------------------------------------------------------------- overlapped gradient all-reduce: bucket logic
def _grad_sync(rank, world):
    from vitok_b200.train import _GradSync
    g = torch.Generator().manual_seed(100 + rank)
    big = [torch.randn(300, 300, generator=g) for _ in range(3)]                    # >= min_numel: async all-reduce each
    strided = torch.randn(300, 600, generator=g)[:, ::2]                            # non-contiguous: goes to the flat bucket
    small = {f"s{i}": torch.randn(17 + i, generator=g) for i in range(4)}           # small tensors: one flat bucket at the end
    local = [t.clone() for t in big] + [strided.clone()] + [t.clone() for t in small.values()]
    sync = _GradSync(None, min_numel=1 << 12)
    for t in big:
        sync.reduce(t)
        t._vtk_reduced = True
    sync.reduce(strided)
    strided._vtk_reduced = True
    named = {f"b{i}": t for i, t in enumerate(big)}
    named["strided"] = strided
    named.update(small)
    named["none"] = None
    sync.finish(named)
    after = big + [strided] + list(small.values())
    # hand-made mean
    ref = []
    for t in local:
        s = t.clone().contiguous()
        dist.all_reduce(s)
        ref.append(s / world)
    return max(float((a - r).abs().max()) for a, r in zip(after, ref)), len(sync.pending), len(sync.small)


def test_grad_sync_buckets_gloo():
    out = _run2("_grad_sync", 29612)
    for r in (0, 1):
        err, pending, small = out[r]
        assert err < 1e-6 and pending == 0 and small == 0


# ---------------------------------------------------------------- reference arm under a 2-rank environment
def test_reference_arm_rank1_is_silent_and_rank0_prints_json():
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", OMP_NUM_THREADS="4")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        env=dict(env, RANK="0", LOCAL_RANK="0"), capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0, r0.stderr[-500:]
    line = json.loads([l for l in r0.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0 and line["unit"] == "images/s"
    # "reference" = the unmodified reference from baseline/_ref (git-ignored: present where it was installed), else the oracle "port"
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "vitok"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
