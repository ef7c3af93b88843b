#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_pp.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
VTK_GEMM_CL4_TRANS=1 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -p no:cacheprovider -k "transposed" 2>&1 | tail -3
python tools/prof_pp.py 512 2>&1 | head -5
VTK_PATCHIFY_U8=rows python tools/prof_pp.py 512 2>&1 | sed -n 4p
python tools/prof_train_gemm.py > $O/r02_train_gemm_cl2.log 2>&1; cat $O/r02_train_gemm_cl2.log
VTK_GEMM_CL4_TRANS=1 python tools/prof_train_gemm.py > $O/r02_train_gemm_cl4.log 2>&1; cat $O/r02_train_gemm_cl4.log
VTK_GEMM_PROF=1 python tools/prof_train_gemm.py 1 > $O/r02_train_gemm_prof.log 2>&1; grep -v "^\[gemm prof\]   epi" $O/r02_train_gemm_prof.log | tail -24
python -m pytest tests/test_gpu_train.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
