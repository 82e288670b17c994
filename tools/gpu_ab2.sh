#!/bin/bash
# Interleaved A/B of the product build against the experimental builds dronesim_b200/libdronesim_b200.<name>.so on ONE box:
# three alternating rounds of the device-resident bench (boxes differ by ~1 %, consecutive runs on one box by ~0.2 %).
# Usage (under gpurun): bash tools/gpu_ab2.sh
set -u
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-sweep --no-parity"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', '%.4f' % d['ms_per_step'], ' '.join('%s %.4f' % (o['workload'][:11], o['ms_per_step']) for o in d['other_workloads']))"; }
for r in 1 2; do
  $B 2>/dev/null | show main
  for lib in dronesim_b200/libdronesim_b200.*.so; do
    [ -f "$lib" ] || continue
    name=$(basename $lib .so); name=${name#libdronesim_b200.}
    DRONESIM_B200_LIB=$PWD/$lib $B 2>/dev/null | show $name
  done
done
