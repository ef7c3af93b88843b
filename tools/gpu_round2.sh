#!/bin/bash
# Round-2 evidence in one single-GPU session: parity suite, c2 bench, ncu launch list of the bench command, full-section
# captures of the dominant kernels (in-model launches), the HBM-bound pp kernels, and a per-kernel profile of the c5 training step.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh TAG [notest]'.  Outputs under gpurun_out/ as r02_*_TAG.*
cd "$(dirname "$0")/.."
TAG=${1:-a}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r02_pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_$TAG.log
fi
python bench.py --steps 20 --warmup 5 > $O/r02_bench_c2_$TAG.json 2> $O/r02_bench_c2_$TAG.err; echo "bench c2 rc=$?"
python bench.py --workload c3 --steps 10 --warmup 3 > $O/r02_bench_c3_$TAG.json 2> $O/r02_bench_c3_$TAG.err; echo "bench c3 rc=$?"
python bench.py --workload c4 --steps 8 --warmup 3 > $O/r02_bench_c4_$TAG.json 2> $O/r02_bench_c4_$TAG.err; echo "bench c4 rc=$?"
python bench.py --workload c5 --steps 4 --warmup 2 > $O/r02_bench_c5_$TAG.json 2> $O/r02_bench_c5_$TAG.err; echo "bench c5 rc=$?"
python bench.py --quantize --steps 20 --warmup 5 > $O/r02_bench_c2_fp8_$TAG.json 2> $O/r02_bench_c2_fp8_$TAG.err; echo "bench c2 fp8 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref_$TAG.json 2> $O/r02_bench_ref_$TAG.err; echo "bench ref rc=$?"
python tools/prof_attn_bwd.py > $O/r02_attn_bwd_$TAG.log 2>&1; cat $O/r02_attn_bwd_$TAG.log
python - <<PY
import json
d = json.loads([l for l in open('$O/r02_bench_c5_$TAG.json') if l.startswith('{')][0]); print('c5', round(d['value'], 2), round(d['ms_per_step'], 1), round(d['model_frac_of_peak'], 3))
d = json.loads([l for l in open('$O/r02_bench_ref_$TAG.json') if l.startswith('{')][0]); print('ref', round(d['value'], 2), d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
for n in ['c2', 'c3', 'c4', 'c2_fp8']:
    d = json.loads([l for l in open('$O/r02_bench_%s_$TAG.json' % n) if l.startswith('{')][0])
    print(n, round(d['value'], 1), round(d['e2e']['value'], 1), round(d['ms_per_step'], 3), round(d['roofline']['frac'], 3), round(d['model_frac_of_peak'], 3),
          {k: round(v['ms_per_step'], 3) for k, v in d['kernel_breakdown'].items()}, d['clocks'])
PY
python tools/prof_pp.py 64 > $O/r02_pp_b64_$TAG.log 2>&1; python tools/prof_pp.py 512 > $O/r02_pp_b512_$TAG.log 2>&1; cat $O/r02_pp_b512_$TAG.log
python tools/prof_train.py c5 > $O/r02_prof_train_c5_$TAG.log 2>&1; echo "prof_train rc=$?"; head -40 $O/r02_prof_train_c5_$TAG.log
# ncu (never a bench value): launch list of the bench command, then one full capture per kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_c2_$TAG.csv python bench.py --steps 2 --warmup 1 > $O/r02_ncu_launches_$TAG.log 2>&1
python tools/ncu_shares.py $O/r02_launches_c2_$TAG.csv patchify > $O/r02_launch_shares_c2_$TAG.txt 2>&1; head -12 $O/r02_launch_shares_c2_$TAG.txt
NCU="ncu --set full --clock-control none --import-source on"
R=/tmp/vtk_ncu; mkdir -p $R      # the .ncu-rep files stay on the box (gpurun copies back at most 64 MiB): only their summaries travel
$NCU -k regex:gemm2_kernel -s 12 -c 1 -f -o $R/r02_prof_qkv_c2_$TAG python tools/prof_step.py c2 1 > $O/r02_ncu_qkv_c2_$TAG.log 2>&1
$NCU -k regex:gemm2_kernel -s 13 -c 1 -f -o $R/r02_prof_resid_c2_$TAG python tools/prof_step.py c2 1 > $O/r02_ncu_resid_c2_$TAG.log 2>&1
$NCU -k regex:attn_persist -s 6 -c 1 -f -o $R/r02_prof_attn_c2_$TAG python tools/prof_step.py c2 1 > $O/r02_ncu_attn_c2_$TAG.log 2>&1
$NCU -k regex:patchify_kernel -c 1 -f -o $R/r02_prof_patchify_$TAG python tools/prof_pp.py 512 > $O/r02_ncu_pp_$TAG.log 2>&1
$NCU -k regex:patchify_u8 -c 1 -f -o $R/r02_prof_patchify_u8_$TAG python tools/prof_pp.py 512 > /dev/null 2>&1
$NCU -k regex:attn_bwd -s 2 -c 2 -f -o $R/r02_prof_attn_bwd_$TAG python tools/prof_attn_bwd.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"patchify|pack_rows|rmsnorm" --csv --log-file $O/r02_hbm_kernels_$TAG.csv python tools/prof_pp.py 512 > /dev/null 2>&1
$NCU -k regex:unpatchify_rows -c 2 -f -o $R/r02_prof_unpatchify_$TAG python tools/prof_pp.py 512 > /dev/null 2>&1
$NCU -k regex:pack_rows_kernel -s 1 -c 1 -f -o $R/r02_prof_pack_rows_$TAG python tools/prof_pp.py 512 > /dev/null 2>&1
for k in qkv_c2 resid_c2 attn_c2 patchify patchify_u8 unpatchify pack_rows attn_bwd; do
  [ -f $R/r02_prof_${k}_$TAG.ncu-rep ] && python tools/ncu_summary.py $R/r02_prof_${k}_$TAG.ncu-rep > $O/r02_ncu_${k}_$TAG.txt 2>&1
done
ls -la $O/*_$TAG.* | cut -c30-; du -sh $O
