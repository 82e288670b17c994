"""Static SASS statistics of the built library -> profiles/r02_sass_stats.txt:
    python tools/sass_report.py > profiles/r02_sass_stats.txt
Library-wide counts of the instructions the design rests on (packed FP32, TMA bulk copies, mbarriers), per-kernel tables
and hot-loop histograms (tools/sass_loop_stats.py) of the step-kernel variants bench.py launches."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "dronesim_b200", "libdronesim_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("# Static SASS statistics of dronesim_b200/libdronesim_b200.so (cuobjdump -sass), round 2 final build")
print("# cubin architectures:", sorted(set(re.findall(r"arch = (sm_\w+)", out))))
tot, perk, cur = collections.Counter(), {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        perk[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(2).split(".")[0]
        tot[op] += 1
        perk[cur][op] += 1
print("# kernels:", len(perk), " total static instructions:", sum(tot.values()))
keys = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "UBLKCP", "UBLKPF", "SYNCS", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG",
        "ATOMG", "DFMA", "LDL", "STL", "UTMALDG"]
print("# library-wide counts of the instructions the design rests on")
for k in keys:
    print("%-8s %d" % (k, tot.get(k, 0)))
print("# FFMA2/FMUL2/FADD2 = packed FP32 (fma/mul/add.f32x2); UBLKCP = cp.async.bulk global->shared (TMA unit, non-tensor 1-D);")
print("# UBLKPF = cp.async.bulk.prefetch.L2; SYNCS = mbarrier operations; no UTMALDG (no tensor-map copies: the tiles are 1-D slabs)")
print()
print("# the step-kernel instantiations the bench launches: hetero16 (mixed types, symmetric downwash, compile-time ground + drag, CoM")
print("# offsets), quad_k8 / traj_quad (homogeneous quads), hexa_circle (homogeneous hexas, CoM offsets)")
names = subprocess.run(["c++filt"], input="\n".join(perk.keys()), capture_output=True, text=True).stdout.splitlines()
want = ("<0, 2, true, true, 0, 3, false, false, true>", "<0, 0, false, true, 0, 0, false, true, false>",
        "<0, 0, true, true, 0, 3, false, true, true>")
for k, n in zip(perk.keys(), names):
    if "ds_step_kernel" in n and any(t in n for t in want):
        c = perk[k]
        print(n)
        print("   total=%d  " % sum(c.values()) + " ".join("%s=%d" % (x, c[x]) for x in keys if c.get(x)))
        sys.stdout.flush()
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loop_stats.py"), k, lib], capture_output=True, text=True).stdout
        print("\n".join("   " + l.strip()[:400] for l in r.splitlines()[1:]))
