"""Print the headline numbers and the per-kernel table of bench.py JSON lines / ncu launch lists under gpurun_out/."""
import csv, json, sys
for f in sys.argv[1:]:
    if f.endswith(".csv"):
        rows = [r for r in csv.reader(open(f)) if len(r) > 5]
        hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
        names = [(r[ki].replace("nvp::<unnamed>::", "")[:50], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
        k = max(i for i, nm in enumerate(names) if "set_gscale" in nm[0])
        tot = 0.0
        for nm, us in names[k:]:
            print(f"   {us:9.1f} us  {nm}"); tot += us
        print(f"   {tot:9.1f} us  total (last step)")
        continue
    lines = [l for l in open(f).read().splitlines() if l.startswith("{")]
    if not lines:
        print(f, "no json"); continue
    d = json.loads(lines[-1])
    print(f, f"value {d['value']:.1f} ms {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f} opt {d.get('with_optimizer', {}).get('value', 0):.1f}")
    for k, v in d.get("kernels", {}).items():
        print(f"   {k:14s} {v['ms_per_step']:.4f} ms  frac {v.get('frac', 0):.3f}")
