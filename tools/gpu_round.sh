#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, then the ncu launch list of the same bench command.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [envs]
set -u
TAG=${1:-r1}
ENVS=${2:-262144}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${TAG}.log
python bench.py --steps 200 --warmup 20 --envs ${ENVS} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
