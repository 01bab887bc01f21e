#!/usr/bin/env python
"""profiles/executed.json + profiles/traffic.json from the round's ncu summaries (tools/ncu_summary.py output under profiles/).
usage: python tools/make_executed.py r02"""
import json, os, re, sys
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
# workload -> (summary file stem, units (segments) per profiled launch)
# (adaptive kernels: attempted steps per segment of the profiled batch, bench.py roofline.attempted_steps_per_segment of the same workload)
MAP = {"direct7_fixed": ("k_direct_cw", 65536, None), "direct6_fixed": ("k_direct_cw_n6", 65536, None), "indirect12": ("k_indirect_cw", 131072, 7.05),
       "indirect14": ("k_indirect_cw14", 131072, 6.94), "indirect12_hc": ("k_indirect_hc", 131072, 7.05), "indirect12_wl": ("k_indirect_wl", 131072, 7.05)}
ex, tr = {}, {}
for wl, (stem, units, att) in MAP.items():
    f = os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (tag, stem))
    if not os.path.exists(f):
        continue
    t = open(f).read()
    num = lambda k: float(re.search(re.escape(k) + r"\s+([0-9.e+]+)", t).group(1))
    unit = lambda k: re.search(re.escape(k) + r"\s+[0-9.e+]+\s+(\S+)", t).group(1)
    fl = num("executed_fp64_flops (2 DFMA + DMUL + DADD, thread level)")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    by = sum(num(k) * scale[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    ex[wl] = {"flops_per_unit": round(fl / units, 1), "flops_per_launch": fl, "units_per_launch": units,
              "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              "source": "profiles/%s_%s_ncu_summary.txt (ncu --set full, one launch; thread-level 2 DFMA + DMUL + DADD)" % (tag, stem)}
    if att:
        ex[wl]["attempted_steps_per_unit"] = att
    tr[wl] = int(by)
ex["_comment"] = ("FP64 work actually executed by the dominant kernel of each workload, from the committed ncu captures; bench.py reports it as "
                  "roofline.executed next to the algorithmic count (profiles/flops_per_unit.json)")
tr["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (same captures); bench.py roofline.traffic"
json.dump(ex, open(os.path.join(ROOT, "profiles", "executed.json"), "w"), indent=1, sort_keys=True)
old = {}
try:
    old = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass
old.update(tr)
json.dump(old, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(ex, indent=1)); print(json.dumps(old, indent=1))
