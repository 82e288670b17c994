#!/bin/bash
# GPU check of a kernel change: parity tests, then the device-resident bench of the product build and of any
# experimental builds dronesim_b200/libdronesim_b200.<name>.so (tools/build_variant.py).
# Usage (under gpurun): bash tools/gpu_ab.sh <tag>
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_${TAG}.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
run() {  # name
  timeout 600 $B > gpurun_out/bench_${TAG}_$1.json 2> gpurun_out/bench_${TAG}_$1.err; echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$1.json")); print("$1 ms/step %.4f  value %.4g  sane %s" % (d["ms_per_step"], d["value"], d["sane"]))
    for o in d["other_workloads"]: print("   %-12s ms %.4f hbm frac %.3f sane %s" % (o["workload"], o["ms_per_step"], o["roofline"]["frac"], o["sane"]))
except Exception as e: print("$1 failed", e)
PY
  tail -2 gpurun_out/bench_${TAG}_$1.err
}
run main
for lib in dronesim_b200/libdronesim_b200.*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so); name=${name#libdronesim_b200.}
  DRONESIM_B200_LIB=$PWD/$lib run var_$name
done
