"""Aggregate an ncu source-page export (cuda,sass) into executed instructions per CUDA source line.
Usage: python tools/ncu_lines.py <report.ncu-rep> [top_n] [samples]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
by_samples = len(sys.argv) > 3 and sys.argv[3] == "samples"  # sort by stall samples instead of executed instructions
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, nk = None, 0
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        continue
    if r[0].isdigit():
        try:
            inst, samp = int(r[7]), int(r[6])
        except Exception:
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += inst
        a[1] += samp
        a[2] = r[1][:100]
tot = sum(v[0] for v in agg.values())
stot = max(1, sum(v[1] for v in agg.values()))
byfile = collections.defaultdict(int)
for (f, l), v in agg.items():
    byfile[f] += v[0]
print("total warp instructions (all captured launches):", tot)
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items() if v})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    print("%-16s %4d inst %5.2f%% samples %5.2f%%  %s" % (f, l, 100 * v[0] / tot, 100 * v[1] / stot, v[2]))
