#!/bin/bash
# Quick GPU check of a kernel change: parity tests + the device-resident bench line only.
# Usage (under gpurun): bash tools/gpu_quick.sh <tag> [extra lib paths to bench as variants ...]
set -u
TAG=${1:-q}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-others"
$B > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); print("main ms/step %.4f  value %.4g  sane %s" % (d["ms_per_step"], d["value"], d["sane"]))
PY
for LIB in "$@"; do
  N=$(basename $LIB .so)
  DRONESIM_B200_LIB=$LIB $B > gpurun_out/bench_${TAG}_${N}.json 2> gpurun_out/bench_${TAG}_${N}.err; echo "$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${N}.json")); print("${N} ms/step %.4f  value %.4g  sane %s" % (d["ms_per_step"], d["value"], d["sane"]))
except Exception as e: print("${N} failed", e)
PY
done
