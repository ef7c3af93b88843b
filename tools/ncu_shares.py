"""Per-kernel-name time shares from an `ncu --csv --metrics gpu__time_duration.sum` log (optionally only launches
with ID >= first_id, to skip set-up/warm-up)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
first = 0
if len(sys.argv) > 2:      # an ID, or a kernel-name substring: start at its first launch (skips model set-up)
    if sys.argv[2].isdigit():
        first = int(sys.argv[2])
    else:
        hdr0 = rows[0]
        k0, i0 = hdr0.index("Kernel Name"), hdr0.index("ID")
        hits = [int(r[i0]) for r in rows[1:] if sys.argv[2] in r[k0]]
        first = min(hits) if hits else 0
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
agg = {}
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum" or int(r[ii]) < first:
        continue
    k = r[ki][:100]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:5d}  {k}")
print(f"total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
