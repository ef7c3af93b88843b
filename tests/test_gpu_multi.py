"""Multi-GPU GPU tests (skipped on a one-GPU box): the data-parallel training step under torch DDP (scripts/train_vae.py:172)
and under the overlapped per-block all-reduce (vitok_b200.enable_grad_sync), launched the way the driver launches bench.py:
one process per GPU through torch.distributed.run on 127.0.0.1, NCCL backend."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, port=29541, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_ddp_and_overlapped_grad_sync_two_ranks():
    """tools/ddp_train_check.py on 2 GPUs: under DDP and under enable_grad_sync every rank ends up with bit-identical gradients
    equal to the mean of the per-rank local gradients, parameters stay in lock-step through FusedAdamW steps, the loss falls."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, "tools/ddp_train_check.py")
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "ddp_train_check OK" in r.stdout and "grad_sync (overlapped all-reduce) OK" in r.stdout


def test_bench_strong_scaling_line_two_ranks():
    """bench.py under torchrun with 2 ranks: c2 switches to strong scaling (64 images dealt 32 per rank, CUDA-graph replay) and
    the line also carries the weak-scaling figure measured in the same run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "5", "--warmup", "3", port=29542)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["config"]["batch_per_gpu"] == 32 and line["config"]["global_batch"] == 64
    assert line["weak"]["scaling"] == "weak" and line["weak"]["batch_per_gpu"] == 64 and line["weak"]["value"] > line["value"] * 0.9
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 32 * 256 * 256 * 3
