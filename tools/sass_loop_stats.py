"""Static SASS statistics of one kernel of the built library and of its innermost hot loop:
    python tools/sass_loop_stats.py <mangled-name-substring> [lib.so]
The hot loop is taken as the longest backward branch span that contains no BAR.SYNC (the K-substep loop of ds_step_kernel).
Counts issued-instruction classes (static, not executed): a check to run before spending GPU time."""
import collections, os, re, subprocess, sys

def main():
    pat = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dronesim_b200", "libdronesim_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(3), m.group(4)))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        print(name, "static instructions:", len(ins))
        best = None
        for addr, op, rest in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < addr:
                        body = [i for i in ins if tgt <= i[0] <= addr]
                        if any(i[1].startswith("BAR") for i in body):
                            continue
                        if best is None or len(body) > len(best):
                            best = body
        for label, body in (("kernel", ins), ("hot loop", best or [])):
            c = collections.Counter()
            for _, op, _ in body:
                k = op.split(".")[0]
                if k == "MUFU":
                    k = op
                c[k] += 1
            fp = sum(c[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
            print("  %-8s n=%d  FMA-pipe=%d  %s" % (label, len(body), fp, " ".join("%s=%d" % kv for kv in c.most_common(40))))

main()
