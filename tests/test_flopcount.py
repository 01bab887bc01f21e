"""CPU: the algorithmic-FLOP figures of bench.py's roofline come from a committed, re-runnable count (tools/flopcount): the kernels'
host-device arithmetic compiled with a counting scalar.  The committed profiles/flops_per_unit.json must be what the tool produces
today, the counted build must compute the same numbers as the plain-double build, and bench.py must read its figures from it."""
import json
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_committed_flop_table_is_reproducible():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "flopcount", "count.py"), "--check"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr


def test_bench_uses_the_counted_table():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "flops_per_unit.json")) as f:
        t = json.load(f)
    assert bench.FLOPS_PER_SEG["direct7"] == t["used"]["direct7_segment"] and bench.FLOPS_PER_SEG["direct6"] == t["used"]["direct6_segment"]
    assert bench.FLOPS_PER_STEP_INDIRECT[12] == t["used"]["indirect12_step"] and bench.FLOPS_PER_STEP_INDIRECT[14] == t["used"]["indirect14_step"]
    for key, survey in t["survey_8d"].items():
        assert t["used"][key] <= survey                      # never more generous than SURVEY.md 8(d)
