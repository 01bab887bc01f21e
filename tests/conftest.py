import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/oracle.py -> oracle/liblto_oracle.so."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden14():
    """The 14-dim system derived symbolically from its Hamiltonian (tests/golden/make_golden14.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "golden14_v1.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def hostcheck():
    """Test-only host build of the kernels' __host__ __device__ arithmetic."""
    d = os.path.join(ROOT, "tests", "native")
    so = os.path.join(d, "liblto_hostcheck.so")
    src = os.path.join(d, "lto_hostcheck.cpp")
    deps = [src] + [os.path.join(ROOT, "lowthrustopt_b200", "csrc", f) for f in ("lto_math.cuh", "lto_prop_generic.cuh", "lto_tableau.h")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def lto():
    """liblto_b200 handle on cuda:0 (GPU tests only)."""
    from lowthrustopt_b200 import capi
    h = capi.Handle(0)
    yield h
    h.close()


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
