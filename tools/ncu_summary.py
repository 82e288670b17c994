"""Print the roofline-relevant metrics of an ncu report (raw page) for each captured launch.
Usage: python tools/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__cycles_elapsed.avg"]
for r in rows[2:]:
    print("=== %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:70], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "", r[hdr.index("Block Size")] if "Block Size" in hdr else ""))
    for w in want:
        if w in hdr:
            print("  %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = [(h.split("issue_stalled_")[1].split("_per")[0], float(r[hdr.index(h)])) for h in hdr
          if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
    print("  stalls per issue:", ", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda kv: -kv[1])[:8]))
