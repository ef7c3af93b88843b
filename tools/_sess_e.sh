#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_pp.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"patchify_u8" -c 6 --csv --log-file $O/r02_hbm_kernels_e.csv python tools/prof_pp.py 512 > /dev/null 2>&1
grep -E "gpu__time_duration|inst_executed|issue_active|dram" $O/r02_hbm_kernels_e.csv | tail -5 | cut -d, -f13-
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"patchify_u8" -c 6 --csv --log-file $O/r02_hbm_kernels_e64.csv python tools/prof_pp.py 64 > /dev/null 2>&1
grep -E "gpu__time_duration" $O/r02_hbm_kernels_e64.csv | tail -2 | cut -d, -f13-
