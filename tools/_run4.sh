timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ae.py -x -q -m gpu > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -4 gpurun_out/r2_pytest4.log
python tools/prof_small.py 4 8 16 2>&1 | tee gpurun_out/r2_small_sk2.log
python tools/prof_graph.py 2>&1 | tee gpurun_out/r2_graph_sk2.log
