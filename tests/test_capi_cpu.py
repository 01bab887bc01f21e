"""CPU: the C-ABI library builds, loads, exports every symbol include/lto_b200.h declares, and
refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


@pytest.fixture(scope="module")
def capi():
    from lowthrustopt_b200 import build, capi as c
    build.build_lib()
    return c


def test_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "lto_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lto_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert set(capi.EXPORTS) == declared
    assert L.lto_version() == 100


def test_struct_layout_matches_header(capi):
    p = capi.direct_params()
    assert (p.MU, p.DU, p.TU) == (0.012150585609624037, 384747.96285603708, 375699.81732246041)   # LowThrustOpt.jl:24-26
    assert p.g0 == 9.81 and p.default_mass == 1000.0 and p.mode == capi.LTO_FIXED
    q = capi.indirect_params()
    assert q.reltol == 1e-13 and q.abstol == 1e-13 and q.err_norm == capi.LTO_NORM_STATE_SENS
    assert C.sizeof(capi.DirectParams) == 7 * 8 + 4 * 4 and C.sizeof(capi.IndirectParams) == 12 * 8 + 4 * 4


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert capi.lib().lto_device_count() == 0
    with pytest.raises(capi.LtoError, match="no CPU fallback"):
        capi.Handle(0)
    from lowthrustopt_b200 import direct
    direct.set_handle(None)
    with pytest.raises(capi.LtoError):
        direct.defectCalc(np.zeros((6, 3)), np.zeros((3, 3)), np.arange(3.0), 6, 3, 10, 2000.0)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (only tests, smoke() and bench.py's cpu legs may)."""
    pkg = os.path.join(ROOT, "lowthrustopt_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, os.path.join(d, f)


def _build_c_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "capi_smoke")
    libdir = os.path.join(ROOT, "lowthrustopt_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "native", "capi_smoke.c"),
                           "-L", libdir, "-llto_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    return exe


def test_header_is_plain_c_and_links(capi, tmp_path):
    """include/lto_b200.h from a C99 translation unit (what cgo / ccall / ctypes bind): compiles with -Wall -Werror, links, and on a
    machine without a GPU the program gets LTO_ERR_NODEVICE from lto_init (no fallback)."""
    import subprocess
    import torch
    exe = _build_c_smoke(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the compute path of the C program is the gpu-marked test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("nodevice:") and "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_program_through_the_abi(capi, tmp_path):
    import subprocess
    out = subprocess.run([_build_c_smoke(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "direct: status 0" in out.stdout and "newton status 0" in out.stdout


def test_host_chunk_plan(capi):
    """The chunk schedule of the host-buffer pipeline (pure host logic): every segment exactly once, whole trajectories, the
    shapes the cost model asks for."""
    # small calls: one shot
    assert capi.host_chunk_plan("direct", 1000) == [1000]
    assert capi.host_chunk_plan("indirect", 29, n_nodes=30) == [29]
    assert capi.host_chunk_plan("direct", 0) == []
    # config 3 (65,536 direct segments, 7 x 20 blocks): copy-bound -> first chunk one wave of the persistent grid, then growing
    plan = capi.host_chunk_plan("direct", 65536, nvar=7)
    assert sum(plan) == 65536 and plan[0] == 148 * 32 and len(plan) <= 6
    assert all(b >= a for a, b in zip(plan[:-2], plan[1:-1])) and all(c % (148 * 32) == 0 for c in plan[:-1])
    # config 4 per GPU (131,072 indirect segments): kernel tail ~ copy time -> a few equal chunks
    plan = capi.host_chunk_plan("indirect", 131072, nvar=12)
    assert sum(plan) == 131072 and 3 <= len(plan) <= 5 and len(set(plan[:-1])) == 1
    plan = capi.host_chunk_plan("indirect", 1 << 20, nvar=12)
    assert sum(plan) == 1 << 20 and 8 <= len(plan) <= 14
    # trajectory forms: chunks are whole trajectories
    for method, nvar, nn, nt in (("indirect", 12, 201, 1024), ("indirect", 14, 201, 777), ("direct", 7, 30, 5000), ("direct", 6, 30, 4097)):
        plan = capi.host_chunk_plan(method, nt * (nn - 1), n_nodes=nn, nvar=nvar)
        assert sum(plan) == nt * (nn - 1) and all(c > 0 and c % (nn - 1) == 0 for c in plan), (method, plan)
    # ragged sizes, both methods, with and without Jacobian, adaptive direct
    rng = np.random.default_rng(5)
    for n in rng.integers(1, 3_000_000, 40):
        for kw in (dict(method="direct", nvar=7), dict(method="direct", nvar=6, jac=False), dict(method="direct", nvar=7, mode=capi.LTO_ADAPTIVE),
                   dict(method="indirect", nvar=12), dict(method="indirect", nvar=14), dict(method="indirect", nvar=12, jac=False)):
            plan = capi.host_chunk_plan(n_seg=int(n), **kw)
            assert sum(plan) == n and min(plan) > 0 and len(plan) <= 300, (n, kw, plan[:4])
    with pytest.raises(capi.LtoError):
        capi.host_chunk_plan("direct", 100, n_nodes=30)        # not a whole number of trajectories


def test_library_carries_sm100a_kernels_only(capi):
    """The shared object holds native sm_100a code for every kernel of the path (no PTX-only JIT path, no other architecture),
    and the FP64 throughput kernels really are DFMA code."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from lowthrustopt_b200 import build
    so = build.build_lib()
    elfs = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    usage = subprocess.run([cuobjdump, "-res-usage", so], capture_output=True, text=True).stdout
    for kern in ("k_direct_cw", "k_direct_state", "k_indirect_cw", "k_indirect_state", "k_indirect_cw14", "k_indirect_state14",
                 "k_indirect_newton", "k_direct_qp", "k_fp64_probe"):
        assert kern in usage, kern
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN3lto3icw16k_indirect_stateENS_12IndirectArgsE", so], capture_output=True, text=True).stdout
    assert sass.count("DFMA") > 500


def test_wrapper_argument_validation():
    """the ctypes wrappers check every auxiliary array before the C side reads n doubles from it -- a scalar
    thrustLimit / rho is broadcast, wrong shapes and unsuitable `out` arrays are ValueErrors (no GPU needed: validation comes first)."""
    from lowthrustopt_b200 import capi
    assert np.array_equal(capi._per_unit("rho", 0.5, 3), [0.5, 0.5, 0.5]) and capi._per_unit("rho", None, 3) is None
    with pytest.raises(ValueError):
        capi._per_unit("thrustLimit", np.ones(2), 3)
    with pytest.raises(ValueError):
        capi._shaped("t0", np.zeros(4), (5,))
    o = {"phi": np.zeros((4, 12, 12), dtype=np.float32)}
    with pytest.raises(ValueError):
        capi._out(o, "phi", (4, 12, 12))
    with pytest.raises(ValueError):
        capi._out({"phi": np.zeros((12, 12, 4)).T}, "phi", (4, 12, 12))          # right shape, not C-contiguous
    assert capi._out({}, "status", (4,), np.int32).dtype == np.int32
