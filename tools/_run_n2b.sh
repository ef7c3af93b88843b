TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_ae.py::test_model_on_second_device_runs_there tests/test_gpu_pp.py -x -q -m gpu 2>&1 | tail -3
python tools/prof_pp.py 512 2>&1 | head -7 | tee gpurun_out/r2_pp_b512_v3.log
run() { name=$1; shift; env "$@" $TR --master-port 29650 bench.py --gpus 2 --workload c5 --steps 5 --warmup 2 > gpurun_out/r2_c5n2_$name.json 2> gpurun_out/r2_c5n2_$name.err
  python - gpurun_out/r2_c5n2_$name.json $name <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 1), d["grad_sync"])
except Exception as ex:
    print(sys.argv[2], "no line", ex)
PY
}
run base X=1
run ch4 NCCL_MAX_NCHANNELS=4
run ch4r8 NCCL_MAX_NCHANNELS=4 VTK_RESERVE_SMS=8
run ch8r16 NCCL_MAX_NCHANNELS=8 VTK_RESERVE_SMS=16
run ch2r4 NCCL_MAX_NCHANNELS=2 VTK_RESERVE_SMS=4
