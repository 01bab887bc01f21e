#!/usr/bin/env python
"""Summarise an .ncu-rep (read offline): headline metrics + stall samples per code region.
usage: ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print("%-80s %s %s" % (k, d[k], u.get(k, "")), file=out)
    try:      # executed FP64 work: thread-level DFMA (2 flops) + DMUL + DADD over the launch (bench.py roofline.executed, profiles/executed.json)
        cyc = float(d["sm__cycles_elapsed.avg"].replace(",", ""))
        g = lambda k: float(d[k].replace(",", "")) * cyc
        fl = 2 * g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") + g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") \
            + g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
        print("%-80s %.6e flop per launch" % ("executed_fp64_flops (2 DFMA + DMUL + DADD, thread level)", fl), file=out)
        for k in ("sm__icc_request_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"):
            if k in d:
                print("%-80s %s" % (k, d[k]), file=out)
    except Exception:
        pass
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[h] or 0) > 0.05:
            print("%-80s %s" % (h, d[h]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
agg = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    k = int(f(r, "Instructions Executed"))
    a = agg.setdefault(k, dict(n=0, samples=0, mix=Counter(), stalls=Counter()))
    a["n"] += 1; a["samples"] += f(r, "# Samples")
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    a["mix"][op[0].split(".")[0] if op else "?"] += 1
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h: a["stalls"][h] += f(r, h)
print("\ncode regions by per-instruction execution count (warp-level):", file=out)
for k, a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:6]:
    print("exec=%d  n_instr=%d  samples=%d" % (k, a["n"], a["samples"]), file=out)
    print("   mix   :", a["mix"].most_common(8), file=out)
    print("   stalls:", [(s, int(v)) for s, v in a["stalls"].most_common(7)], file=out)
