#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_pp.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"patchify_u8" -c 6 --csv --log-file $O/r02_hbm_kernels_d.csv python tools/prof_pp.py 512 > /dev/null 2>&1
grep -E "gpu__time_duration|inst_executed" $O/r02_hbm_kernels_d.csv | tail -4 | cut -d, -f5,13-
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"patchify_u8" -c 6 --csv --log-file $O/r02_hbm_kernels_d64.csv python tools/prof_pp.py 64 > /dev/null 2>&1
grep -E "gpu__time_duration" $O/r02_hbm_kernels_d64.csv | tail -2 | cut -d, -f5,13-
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r02_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_d.log
