#!/bin/bash
# One GPU session for the round's evidence: parity tests, bench (c2, c3, c4, reference arm), ncu launch list of the bench
# command and full-section captures of the dominant kernels.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
TAG=${1:-v5}
mkdir -p gpurun_out
timeout 900 bash tools/gpu_check.sh tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_pp.py tests/test_gpu_attention.py tests/test_gpu_packing.py tests/test_gpu_ae.py tests/test_gpu_fp8.py tests/test_gpu_train.py 2>&1 | grep -E "^===|passed|failed|FAILED|ERROR"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"
python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench c3 rc=$?"
python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; echo "bench c4 rc=$?"
python bench.py --workload c5 --steps 3 --warmup 2 > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "bench c5 rc=$?"
python bench.py --quantize --steps 20 --warmup 5 > gpurun_out/bench_c2_fp8_$TAG.json 2> gpurun_out/bench_c2_fp8_$TAG.err; echo "bench c2 fp8 rc=$?"
python bench.py --quantize --workload c4 --steps 8 --warmup 3 > gpurun_out/bench_c4_fp8_$TAG.json 2> gpurun_out/bench_c4_fp8_$TAG.err; echo "bench c4 fp8 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
python -c "
import json
for n in ['c2','c3','c4','c2_fp8','c4_fp8']:
    d=json.loads([l for l in open('gpurun_out/bench_%s_$TAG.json' % n) if l.startswith('{')][0]); print(n, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['model_frac_of_peak'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()}, d['clocks'])
d=json.loads([l for l in open('gpurun_out/bench_c5_$TAG.json') if l.startswith('{')][0]); print('c5', round(d['value'],2), round(d['ms_per_step'],1), round(d['model_frac_of_peak'],3))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c2_$TAG.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/ncu_shares.py gpurun_out/launches_c2_$TAG.csv patchify > gpurun_out/launch_shares_c2_$TAG.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
# in-model kernels of the first encode+decode (tools/prof_step.py): gemm2 launch 12 = decoder block 1 QKV+fc1, 13 = its out_proj+fc2
$NCU -k regex:gemm2_kernel -s 12 -c 1 -f -o gpurun_out/prof_qkv_c2_$TAG python tools/prof_step.py c2 1 > gpurun_out/ncu_qkv_c2_$TAG.log 2>&1
$NCU -k regex:gemm2_kernel -s 13 -c 1 -f -o gpurun_out/prof_resid_c2_$TAG python tools/prof_step.py c2 1 > gpurun_out/ncu_resid_c2_$TAG.log 2>&1
$NCU -k regex:attn_persist -s 6 -c 1 -f -o gpurun_out/prof_attn_c2_$TAG python tools/prof_step.py c2 1 > gpurun_out/ncu_attn_c2_$TAG.log 2>&1
$NCU -k regex:attn_kernel -s 6 -c 1 -f -o gpurun_out/prof_attn_c4_$TAG python tools/prof_step.py c4 1 > gpurun_out/ncu_attn_c4_$TAG.log 2>&1
$NCU -k regex:patchify -c 1 -f -o gpurun_out/prof_patchify_$TAG python tools/prof_pp.py > gpurun_out/ncu_pp_$TAG.log 2>&1
$NCU -k regex:unpatchify_kernel -c 1 -f -o gpurun_out/prof_unpatchify_$TAG python tools/prof_pp.py > /dev/null 2>&1
$NCU -k regex:rmsnorm -c 1 -f -o gpurun_out/prof_rmsnorm_$TAG python tools/prof_pp.py > /dev/null 2>&1
$NCU -k regex:pack_rows_kernel -s 1 -c 1 -f -o gpurun_out/prof_pack_rows_$TAG python tools/prof_pp.py > /dev/null 2>&1
ls -la gpurun_out/*$TAG*
