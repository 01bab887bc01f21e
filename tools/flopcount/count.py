"""Builds tools/flopcount/flopcount.cpp (the kernels' host-device arithmetic with a FLOP-counting scalar) and writes
profiles/flops_per_unit.json -- the counted algorithmic FLOPs per unit bench.py's roofline cites.

    python tools/flopcount/count.py [--check]     (--check: compare with the committed JSON instead of writing it)
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
OUT = os.path.join(ROOT, "profiles", "flops_per_unit.json")

# SURVEY.md 8(d) survey-time figures (first-order minimal formulation, counted by hand / throwaway script)
SURVEY = {"direct7_segment": 284040, "direct6_segment": 217836, "indirect12_step": 39468, "indirect14_step": 46462}


def run(plain=False):
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "flopcount")
        cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, os.path.join(HERE, "flopcount.cpp")]
        if plain:
            cmd.insert(1, "-DLTO_FLOPCOUNT_PLAIN")
        subprocess.check_call(cmd)
        return json.loads(subprocess.check_output([exe], text=True))


def table():
    counted = run()
    plain = run(plain=True)                       # same code in plain double: the counted runs compute the same numbers
    for k in counted:
        a, b = counted[k]["checksum"], plain[k]["checksum"]
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (k, a, b)
    c = {k: v["flops"] for k, v in counted.items()}
    used = {
        "direct7_segment": min(SURVEY["direct7_segment"], c["direct7_segment"]),
        "direct6_segment": min(SURVEY["direct6_segment"], c["direct6_segment"]),
        "indirect12_step": min(SURVEY["indirect12_step"], c["indirect12_step_first_order"], c["indirect12_step_half_column"]),
        "indirect14_step": min(SURVEY["indirect14_step"], c["indirect14_step_first_order"]),
    }
    return {"rule": "add/sub/mul/div/sqrt = 1, fma = 2, exp/tanh/pow = 1 per call; comparisons, fabs, fmax/fmin, negation = 0 (SURVEY.md 8(d))",
            "how": "tools/flopcount/count.py: the kernels' __host__ __device__ arithmetic compiled with a counting scalar "
                   "(tools/flopcount/flopcount.cpp); checksums equal the plain-double build's",
            "counted": counted, "survey_8d": SURVEY,
            "used": used,
            "used_is": "min(survey-time figure, every counted formulation): the roofline's algorithmic FLOPs never exceed what a counted, "
                       "verified formulation performs"}


if __name__ == "__main__":
    t = table()
    if "--check" in sys.argv:
        with open(OUT) as f:
            ref = json.load(f)
        assert ref["used"] == t["used"] and {k: v["flops"] for k, v in ref["counted"].items()} == {k: v["flops"] for k, v in t["counted"].items()}
        print("ok")
    else:
        with open(OUT, "w") as f:
            json.dump(t, f, indent=1)
        print(json.dumps(t["used"]), {k: v["flops"] for k, v in t["counted"].items()})
