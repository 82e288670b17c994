"""Regenerate profiles/roofline_inputs.json from a committed `ncu --set full --import-source on` report of the bench command:

    python tools/roofline_inputs.py profiles/<name>.ncu-rep [--vehicles 4194304] [--out profiles/roofline_inputs.json]

Per launch of the step kernel it extracts
  * executed FP32 FLOP, from the per-SASS-instruction "Predicated-On Thread Instructions Executed" counters of the source
    page, summed per opcode: FFMA 2, FMUL / FADD 1, and the PACKED forms FFMA2 4, FMUL2 / FADD2 2 (ncu's
    smsp__sass_thread_inst_executed_op_f{fma,mul,add}_pred_on metrics do not see the packed opcodes at all);
  * warp instructions issued and the packed-instruction share;
  * DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the launch duration under ncu.
The file is stamped with dronesim_b200._lib.source_hash(); bench.py ignores it when the sources have changed since."""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

FLOP = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2}


def source_blocks(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur, hdr = [], None, None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"kernel": r[1], "rows": []}
            blocks.append(cur)
            hdr = None
        elif r[0] == "Address":
            hdr = r
        elif cur is not None and hdr is not None and len(r) >= len(hdr) - 1 and r[0].startswith("0x"):
            cur["rows"].append(dict(zip(hdr, r)))
    return blocks


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6,
         "msecond": 1e6, "s": 1e9, "second": 1e9}


def raw_rows(rep):
    """Rows of the raw page with byte / time metrics converted to bytes / ns (the page prints them in scaled units)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for k, u in zip(hdr, units):
            if u in SCALE and k in d:
                try:
                    d[k] = str(float(d[k].replace(",", "")) * SCALE[u])
                except ValueError:
                    pass
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--vehicles", type=int, default=4194304)
    ap.add_argument("--kernel", default="ds_step_kernel")
    ap.add_argument("--workload", default="hetero16, 262144 envs x 16 drones, K=8")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "roofline_inputs.json"))
    ap.add_argument("--opcodes", default="", help="also write the per-opcode executed-instruction table (CSV)")
    a = ap.parse_args()
    from dronesim_b200 import _lib

    blocks = [b for b in source_blocks(a.report) if a.kernel in b["kernel"]]
    raws = [r for r in raw_rows(a.report) if a.kernel in r.get("Kernel Name", "")]
    if not blocks or not raws:
        raise SystemExit("no %s launch in %s" % (a.kernel, a.report))
    b, raw = blocks[0], raws[0]
    thr, warp = collections.Counter(), collections.Counter()
    for row in b["rows"]:
        toks = row["Source"].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        thr[op] += int(row["Predicated-On Thread Instructions Executed"])
        warp[op] += int(row["Instructions Executed"])
    flop = sum(thr[k] * v for k, v in FLOP.items())
    fp_warp = sum(warp[k] for k in FLOP)
    packed_warp = sum(warp[k] for k in ("FFMA2", "FMUL2", "FADD2"))

    def num(k):
        return float(raw[k].replace(",", ""))

    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    unit = 1.0
    out = {
        "source_hash": _lib.source_hash(),
        "report": os.path.relpath(a.report, ROOT),
        "kernel": b["kernel"],
        "workload": a.workload,
        "vehicles_per_launch": a.vehicles,
        "fp32_flop_executed_per_launch": float(flop),
        "fp32_thread_inst_per_launch": {k: int(thr[k]) for k in FLOP},
        "warp_inst_per_launch": int(sum(warp.values())),
        "fp32_warp_inst_per_launch": int(fp_warp),
        "packed_fp32_warp_inst_per_launch": int(packed_warp),
        "dram_bytes_per_launch": dram * unit,
        "ncu_duration_ns": num("gpu__time_duration.sum"),
        "ncu": {k: raw.get(k) for k in (
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")},
        "how": "tools/roofline_inputs.py: FLOP = sum over SASS opcodes of predicated-on thread instructions x {FFMA 2, FMUL 1, FADD 1, "
               "FFMA2 4, FMUL2 2, FADD2 2}; first captured launch of the kernel",
    }
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    if a.opcodes:  # the per-opcode table the FLOP count was summed from (committed next to the JSON; the report itself is 10+ MB)
        with open(a.opcodes, "w") as f:
            f.write("opcode,warp_instructions_executed,predicated_on_thread_instructions_executed,flop_per_thread_instruction\n")
            for k, v in sorted(warp.items(), key=lambda kv: -kv[1]):
                f.write("%s,%d,%d,%d\n" % (k, v, thr[k], FLOP.get(k, 0)))
    per_v = flop / a.vehicles
    print(json.dumps({k: out[k] for k in ("source_hash", "kernel", "fp32_flop_executed_per_launch", "warp_inst_per_launch",
                                          "dram_bytes_per_launch", "ncu_duration_ns")}, indent=1))
    print("per vehicle and control step: %.0f FLOP, %.0f thread instructions; FP32 thread instructions %s" % (
        per_v, sum(thr.values()) / a.vehicles, {k: round(thr[k] / a.vehicles, 1) for k in FLOP}))


if __name__ == "__main__":
    main()
