"""Dump the SASS of one kernel from an .ncu-rep source page with stall samples: python scripts/ncu_src.py REP [first] [last] [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0; hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; idx = 0
tot = {}
for r in rows:
    if r and r[0] == "Address": h = r; idx = 0; stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]; continue
    if h is None or len(r) < len(h): continue
    samples = int(r[2] or 0)
    if lo <= idx <= hi and samples >= mins:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:3]
        print(f"#{idx:<5d} {samples:6d} exec {r[5]:>9s}  {r[1].strip()[:64]:64s} " + " ".join(f"{n}:{c}" for c, n in st if c))
    for i in stall_cols: tot[h[i][6:]] = tot.get(h[i][6:], 0) + int(r[i] or 0)
    idx += 1
if mins == 0 and lo == 0 and hi == 10**9: pass
print("totals:", sorted(tot.items(), key=lambda kv: -kv[1]))
