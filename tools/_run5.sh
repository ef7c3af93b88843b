timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -12 gpurun_out/r2_pytest5.log
grep "\[parity\] gpu oracle\|5B-shape\|drop_path train\|non-prefix\|activations held" gpurun_out/r2_pytest5.log
python tools/attn_vs_fa2.py > gpurun_out/r2_attn_vs_fa2.txt 2>&1; cat gpurun_out/r2_attn_vs_fa2.txt | tail -12
python tools/gpu_oracle.py c2 c3 > gpurun_out/r2_gpu_oracle.txt 2>&1; tail -4 gpurun_out/r2_gpu_oracle.txt
python bench.py --workload c5 --steps 5 --warmup 2 > gpurun_out/r2_bench_c5_a.json 2> gpurun_out/r2_bench_c5_a.err; tail -c 900 gpurun_out/r2_bench_c5_a.json; tail -3 gpurun_out/r2_bench_c5_a.err
