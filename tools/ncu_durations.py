"""Compact per-kernel summary of an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*.sum]` log:
median duration (us), launches and median DRAM bytes per kernel name."""
import csv
import statistics
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = {}
for r in rows[1:]:
    per.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = {}
for (_, k), m in sorted(per.items()):
    agg.setdefault(k, []).append(m)
for k, ms in agg.items():
    if any(s in k for s in ("at::native", "elementwise_kernel", "FillFunctor")):
        continue
    t = statistics.median(m.get("gpu__time_duration.sum", 0.0) for m in ms) / 1e3
    rd = statistics.median(m.get("dram__bytes_read.sum", 0.0) for m in ms) / 1e6
    wr = statistics.median(m.get("dram__bytes_write.sum", 0.0) for m in ms) / 1e6
    print(f"{k[:90]:90s} n={len(ms):3d} median {t:8.1f} us  dram rd {rd:7.1f} MB wr {wr:7.1f} MB")
