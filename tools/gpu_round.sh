#!/bin/bash
# One GPU-box visit: the full default bench line (both arms), optionally followed by the ncu evidence for the same build.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [profile]
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
S=$(date +%s)
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.4g ms %.4f  median %.4f p95 %.4f  sane %s" % (d["value"], d["ms_per_step"], d["per_step"]["median_ms"], d["per_step"]["p95_ms"], d["sane"]))
r=d["roofline"]; print("roofline", r["bound"], "%.3f" % r["frac"], r["unit"], "hbm %.3f" % r["hbm"]["frac"], "issue", r.get("issue",{}).get("frac"), "traffic", r["traffic"], r.get("inputs"))
e=d["e2e"]; print("e2e %.4g (%.3f ms) compact %.4g (%.3f ms) sync %.4g obs %.4g" % (e["value"], e["ms_per_step"], e["compact_targets"]["value"], e["compact_targets"]["ms_per_step"], e["per_step_sync"]["value"], e["with_full_obs"]["value"]))
print("parity", d["parity_check"])
c=d["cpu_baseline"]; print("cpu", c["kind"], c.get("detail"), "%.4g" % c["value"], c["cores"], "vectorised", c["vectorised"]["value"])
for o in d["other_workloads"]: print("   %-40s N %9d ms %.4f value %.4g %s sane %s" % (o["workload"], o["vehicles_per_gpu"], o["ms_per_step"], o["value"], ("hbm %.3f" % o["roofline"]["frac"]) if "roofline" in o else "", o["sane"]))
PY
S=$(date +%s)
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "reference arm rc=$? wall $(( $(date +%s) - S )) s"
head -c 600 gpurun_out/bench_${TAG}_ref.json; echo
[ "${2:-}" = "profile" ] && bash tools/gpu_profile.sh ${TAG}
