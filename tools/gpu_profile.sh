#!/bin/bash
# ncu evidence for the round: full-section captures of the block GEMMs and the attention kernel (one launch each),
# plus the sustained-clock check.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v2}
python tools/prof_gemm.py c2 sustain > gpurun_out/sustain_c2_$TAG.log 2>&1
python tools/prof_gemm.py c4 sustain > gpurun_out/sustain_c4_$TAG.log 2>&1
cat gpurun_out/sustain_c2_$TAG.log gpurun_out/sustain_c4_$TAG.log
python tools/prof_attn.py > gpurun_out/attn_plain_$TAG.log 2>&1; cat gpurun_out/attn_plain_$TAG.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_qkv_c2_$TAG python tools/prof_gemm.py c2 > gpurun_out/ncu_qkv_c2_$TAG.log 2>&1
$NCU -k regex:gemm_kernel -s 26 -c 1 -f -o gpurun_out/prof_resid_c2_$TAG python tools/prof_gemm.py c2 > gpurun_out/ncu_resid_c2_$TAG.log 2>&1
$NCU -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_qkv_c4_$TAG python tools/prof_gemm.py c4 > gpurun_out/ncu_qkv_c4_$TAG.log 2>&1
$NCU -k regex:attn_kernel -s 3 -c 1 -f -o gpurun_out/prof_attn_c2_$TAG python tools/prof_attn.py > gpurun_out/ncu_attn_c2_$TAG.log 2>&1
$NCU -k regex:attn_kernel -s 16 -c 1 -f -o gpurun_out/prof_attn_c4_$TAG python tools/prof_attn.py > gpurun_out/ncu_attn_c4_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
