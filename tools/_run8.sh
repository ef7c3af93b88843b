timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
tail -6 gpurun_out/r2_pytest8.log
grep "\[parity\] gpu oracle\|5B-shape\|drop_path train\|non-prefix\|activations held\|SSIM" gpurun_out/r2_pytest8.log
python tools/prof_pp.py 64 2>&1 | tee gpurun_out/r2_pp_b64_v2.log
python tools/prof_pp.py 512 2>&1 | tee gpurun_out/r2_pp_b512_v2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; tail -c 600 gpurun_out/r2_bench_c2_b.json; tail -3 gpurun_out/r2_bench_c2_b.err
python bench.py --workload 350M-2048 --steps 5 --warmup 3 > gpurun_out/r2_bench_2048.json 2> gpurun_out/r2_bench_2048.err; head -c 700 gpurun_out/r2_bench_2048.json; tail -3 gpurun_out/r2_bench_2048.err
python bench.py --workload 350M-4096 --sw 4096 --steps 3 --warmup 3 > gpurun_out/r2_bench_4096sw.json 2> gpurun_out/r2_bench_4096sw.err; head -c 700 gpurun_out/r2_bench_4096sw.json; tail -3 gpurun_out/r2_bench_4096sw.err
