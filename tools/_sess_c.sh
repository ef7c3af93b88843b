#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"patchify|pack_rows|rmsnorm" --csv --log-file $O/r02_hbm_kernels_c.csv python tools/prof_pp.py 512 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_hbm_kernels_c.csv')) if len(r)>10]
h=rows[0]; ki,mi,vi,ii=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki][:70]),{})[r[mi]]=float(r[vi].replace(',',''))
seen={}
for (i,k),m in sorted(d.items()):
    seen.setdefault(k,[]).append(m)
for k,l in seen.items():
    m=l[len(l)//2]
    print(f"{k:72s} n={len(l):3d} {m}")
PY
ncu --set full --clock-control none --import-source on -k regex:patchify_u8_strip -s 2 -c 1 -f -o $O/r02_prof_patchify_strip_c python tools/prof_pp.py 512 > /dev/null 2>&1
python tools/ncu_summary.py $O/r02_prof_patchify_strip_c.ncu-rep > $O/r02_ncu_patchify_strip_c.txt 2>&1
grep -E "gpu__time_duration.sum|dram__bytes|lsu_wavefronts.sum.pct|issue_active|warps_active|stalled_(long|short|barrier|membar|lg|mio|math|wait|no_inst|drain)" $O/r02_ncu_patchify_strip_c.txt | cut -c1-70,100-170
for e in 0 1; do VTK_GEMM_CL4_TRANS=$e python bench.py --workload c5 --steps 4 --warmup 2 > $O/r02_bench_c5_cl4t$e.json 2> $O/r02_bench_c5_cl4t$e.err; python -c "
import json; d=json.loads([l for l in open('$O/r02_bench_c5_cl4t$e.json') if l.startswith('{')][0]); print('c5 CL4_TRANS=$e', d['ms_per_step'], d['clocks'])"; done
