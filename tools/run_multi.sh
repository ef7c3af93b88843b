#!/bin/bash
# Multi-GPU evidence of one box (BASELINE configs 2, 4, 5 at N GPUs): usage  gpurun --gpus N -- 'bash tools/run_multi.sh N [what...]'
# what = c2 c4 c5 c5r c5ddp test (default: c2 c4 c5 c5ddp test); c5r = c5 with 16 SMs reserved for NCCL (8 channels).  Lines land in gpurun_out/r02_bench_<workload>_n<N>*.json.
N=${1:-2}; shift
WHAT=${@:-c2 c4 c5 c5ddp test}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
for w in $WHAT; do
  case $w in
    c2)    $TR --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err ;;
    c4)    $TR --master-port 29602 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err ;;
    c5)    $TR --master-port 29603 bench.py --gpus $N --workload c5 --steps 6 --warmup 2 > gpurun_out/r02_bench_c5_n${N}_sync.json 2> gpurun_out/r02_bench_c5_n${N}_sync.err ;;
    c5r)   NCCL_MAX_NCHANNELS=8 VTK_RESERVE_SMS=16 $TR --master-port 29605 bench.py --gpus $N --workload c5 --steps 6 --warmup 2 > gpurun_out/r02_bench_c5_n${N}_sync_reserve16.json 2> gpurun_out/r02_bench_c5_n${N}_sync_reserve16.err ;;
    c5ddp) VTK_TRAIN_DDP=1 $TR --master-port 29604 bench.py --gpus $N --workload c5 --steps 6 --warmup 2 > gpurun_out/r02_bench_c5_n${N}_ddp.json 2> gpurun_out/r02_bench_c5_n${N}_ddp.err ;;
    test)  python -m pytest tests/test_gpu_multi.py tests/test_gpu_ae.py::test_model_on_second_device_runs_there -x -q -m gpu > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/r02_pytest_multi_n$N.log ;;
  esac
  echo "== $w rc=$?"
done
for f in gpurun_out/r02_bench_*_n$N*.json; do echo "--- $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    keep = {k: d.get(k) for k in ("metric", "value", "ms_per_step", "n_gpus", "scaling", "model_frac_of_peak", "grad_sync", "peak_mem_gb")}
    keep["e2e"] = d.get("e2e", {}).get("value")
    keep["weak"] = (d.get("weak") or {}).get("value")
    keep["strong"] = (d.get("strong") or {}).get("value")
    print(keep)
except Exception as ex:
    print("no line:", ex)
PY
done
tail -3 gpurun_out/r02_bench_*_n$N*.err 2>/dev/null | grep -v "^$" | tail -20
