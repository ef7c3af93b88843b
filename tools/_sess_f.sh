#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_train.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
python tools/prof_train.py c5 > $O/r02_prof_train_c5_f.log 2>&1; echo "prof_train rc=$?"; grep -v Warn $O/r02_prof_train_c5_f.log | head -24
python bench.py --workload c5 --steps 4 --warmup 2 > $O/r02_bench_c5_f.json 2> $O/r02_bench_c5_f.err; python -c "
import json; d=json.loads([l for l in open('$O/r02_bench_c5_f.json') if l.startswith('{')][0]); print('c5', d['ms_per_step'], d['model_frac_of_peak'], d['clocks'])"
