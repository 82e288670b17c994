#!/bin/bash
# Time every experimental build  dronesim_b200/libdronesim_b200.<name>.so  with the headline bench (device-resident only).
mkdir -p gpurun_out
for lib in dronesim_b200/libdronesim_b200.*.so; do
  name=$(basename $lib .so); name=${name#libdronesim_b200.}
  DRONESIM_B200_LIB=$PWD/$lib python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_$name.json"))
    print("%-14s ms_per_step %.4f  value %.3e  sane %s" % ("$name", d["ms_per_step"], d["value"], d["sane"]))
except Exception as e:
    print("$name", "failed", e)
PY
done
