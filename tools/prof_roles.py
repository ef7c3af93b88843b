"""VTK_GEMM_PROF=1 driver: one launch of each block GEMM at the c2 / c4 shapes; the library prints per-role cycles."""
import os
import subprocess
import sys

os.environ["VTK_GEMM_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for wl in sys.argv[1:] or ["c2", "c4"]:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_gemm.py"), wl, "once"], capture_output=True, text=True)
    lines = [l for l in r.stderr.splitlines() if "gemm prof" in l]
    # after 3 warm-ups the 4th report of each kind is a steady-state launch
    print(f"--- {wl}")
    seen = {}
    for i in range(0, len(lines) - 1, 2):
        kind = lines[i].split("]")[1].split(":")[0].strip()
        seen.setdefault(kind, []).append((lines[i], lines[i + 1]))
    for kind, v in seen.items():
        a, b = v[min(5, len(v) - 1)]
        print(a)
        print(b)
