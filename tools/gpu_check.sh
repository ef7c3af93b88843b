#!/bin/bash
# Run every GPU parity test file in its own process (a device-side trap poisons the CUDA context of the
# process that hit it) with a hard timeout, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_pp.py tests/test_gpu_attention.py tests/test_gpu_ae.py"}
rc=0
for f in $FILES; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout 600 python -m pytest $f -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  [ $r -ne 0 ] && rc=$r
  grep -E "^\[parity\]|^\[probe\]|^(PASSED|FAILED|ERROR)|passed|failed|vtk:" gpurun_out/$name.log | tail -60
done
exit $rc
