#!/bin/bash
# Multi-GPU visit (gpurun --gpus 8): concurrent H2D bandwidth at 1/2/4/8 ranks, then the bench at 8, 4, 2 GPUs.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for G in ${H2D_RANKS-1 2 4 8}; do   # H2D_RANKS="" skips the bandwidth micro-benchmark
  $TR --nproc-per-node $G --master-port $((29500+G)) tools/ubench/h2d_concurrent.py 2>/dev/null | grep ranks | tee -a gpurun_out/h2d_concurrent_${TAG}.txt
done
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/topo_${TAG}.txt
for G in ${BENCH_RANKS-8 4 2}; do
  $TR --nproc-per-node $G --master-port $((29600+G)) bench.py --gpus $G --steps 100 --warmup 10 --no-others > gpurun_out/bench_${G}gpu_${TAG}.json 2> gpurun_out/bench_${G}gpu_${TAG}.err
  echo "bench $G GPUs rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${G}gpu_${TAG}.json")); e=d["e2e"]
    print("  N=%d value %.4g ms %.4f | e2e %.4g (%.3f ms, %.1f GB/s per GPU) | compact %.4g (%.3f ms) | sane %s" % (d["n_gpus"], d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e["compact_targets"]["value"], e["compact_targets"]["ms_per_step"], d["sane"]))
except Exception as ex: print("  failed", ex)
PY
done
