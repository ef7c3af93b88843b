timeout 1500 python -m pytest tests/test_gpu_packing.py tests/test_gpu_pp.py tests/test_gpu_train.py tests/test_gpu_reference.py tests/test_gpu_multi.py tests/test_gpu_fp8.py -x -q -m gpu > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -12 gpurun_out/r2_pytest6.log
grep "\[parity\] gpu oracle\|5B-shape\|drop_path train\|non-prefix\|activations held" gpurun_out/r2_pytest6.log
