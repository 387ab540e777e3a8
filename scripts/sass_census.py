"""SASS census of the built library: per kernel, how many tcgen05 / TMEM / TMA / mbarrier instructions it contains
(cuobjdump -sass; B200_PROFILING.md "What proves a Blackwell-native kernel").  usage: python scripts/sass_census.py [LIB]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "nvp_b200/libnvp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"UTCHMMA (tcgen05.mma kind::f16)": r"\bUTCHMMA\b", "LDTM (tcgen05.ld)": r"\bLDTM", "UTCBAR (tcgen05.commit)": r"\bUTCBAR",
        "UBLKCP (cp.async.bulk)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "LDGSTS (cp.async)": r"\bLDGSTS", "MUFU.SIN/COS": r"MUFU\.(SIN|COS)",
        "RED (red.global)": r"\bRED\b|\bREDG\b", "HMMA (legacy mma.sync)": r"\bHMMA\b", "instructions": r"^\s+/\*[0-9a-f]{4,}\*/"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"nvp::\(anonymous namespace\)::|\(anonymous namespace\)::", "", name)[:110]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {lib}: cuobjdump -sass census (instruction counts per kernel; static, not executed counts)")
keys = list(pats)
for name, c in counts.items():
    if c["instructions"] == 0:
        continue
    print(f"{name}\n    " + "  ".join(f"{k.split(' ')[0]}={c[k]}" for k in keys if c[k]))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("TOTAL  " + "  ".join(f"{k.split(' ')[0]}={tot[k]}" for k in keys))
