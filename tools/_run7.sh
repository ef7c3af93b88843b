timeout 1500 python -m pytest tests/test_gpu_pp.py tests/test_gpu_train.py tests/test_gpu_reference.py tests/test_gpu_multi.py tests/test_gpu_fp8.py -x -q -m gpu > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -6 gpurun_out/r2_pytest7.log
grep "\[parity\] gpu oracle\|5B-shape\|drop_path train\|non-prefix\|activations held" gpurun_out/r2_pytest7.log
python tools/prof_pp.py 64 2>&1 | tee gpurun_out/r2_pp_b64.log
python tools/prof_pp.py 512 2>&1 | tee gpurun_out/r2_pp_b512.log
