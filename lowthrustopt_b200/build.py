"""Builds liblto_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m lowthrustopt_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels with a repo snapshot.
There is deliberately no other backend: if nvcc or the sources are missing this raises.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(HERE, "..", "include")
OBJ = os.environ.get("LTO_OBJ_DIR") or os.path.join(HERE, "_build")
LIB = os.environ.get("LTO_LIB_OUT") or os.path.join(HERE, "liblto_b200.so")     # (LTO_LIB_OUT, LTO_EXTRA_SOURCES: tools/experiments/build_variant.sh)
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v", "-ccbin", "/usr/bin/g++"] + os.environ.get("LTO_NVCC_EXTRA", "").split()   # (development: -D switches)


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: liblto_b200.so cannot be built (there is no non-CUDA backend)")
    return nvcc


def _sources():
    extra = [os.path.abspath(f) for f in os.environ.get("LTO_EXTRA_SOURCES", "").split(":") if f]
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu")) + extra


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(INC, "lto_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + ARCH + NVCC_FLAGS + ["-I", INC, "-I", CSRC, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        return r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(compile_one, jobs):
                if verbose:
                    sys.stderr.write(out)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in _sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
