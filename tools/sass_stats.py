"""Per-kernel SASS statistics of the built library:  python tools/sass_stats.py [substring ...]
Counts issued-instruction classes per kernel (static, not executed) - a quick check before spending GPU time."""
import re, subprocess, sys, os
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dronesim_b200", "libdronesim_b200.so")
if len(sys.argv) > 1 and sys.argv[1].endswith(".so"):
    lib = sys.argv.pop(1)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, stats = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(2).split(".")[0]
        if op == "MUFU": op = m.group(2)
        stats[cur][op] = stats[cur].get(op, 0) + 1
pats = sys.argv[1:]
for k, v in stats.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    if pats and not any(p in name for p in pats):
        continue
    tot = sum(v.values())
    top = sorted(v.items(), key=lambda kv: -kv[1])[:14]
    print("%s  total=%d  %s" % (name, tot, " ".join("%s=%d" % kv for kv in top)))
    print("    MUFU:", {a: b for a, b in v.items() if a.startswith("MUFU")}, "SHFL=%d LDS=%d STS=%d LDL=%d STL=%d" % tuple(v.get(x, 0) for x in ("SHFL", "LDS", "STS", "LDL", "STL")))
