import os
import sys

import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the oracle (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_api
    return oracle_api.load()


@pytest.fixture(scope="session")
def P(built):
    import pbrt_r3_b200
    return pbrt_r3_b200


BUILD = os.path.join(ROOT, "tests", "hostcheck", "_build")
HOSTCHECK_LIB = os.path.join(BUILD, "libpbrt_hostcheck.so")


def build_hostcheck():
    """tests/hostcheck: the product's CUDA sources compiled by g++ against a serial stand-in for the CUDA runtime (TEST
    INFRASTRUCTURE; the product library never contains this build)."""
    os.makedirs(BUILD, exist_ok=True)
    csrc = os.path.join(ROOT, "pbrt-r3_b200", "csrc")
    srcs = [os.path.join(csrc, "cuda", "device_api.cu")] + [os.path.join(csrc, "host", f) for f in ("front.cpp", "api_common.cpp", "bvh_build.cpp", "scene_accel.cpp")]
    deps = srcs + [os.path.join(csrc, "cuda", f) for f in os.listdir(os.path.join(csrc, "cuda")) if f.endswith((".cuh", ".cu"))]
    deps += [os.path.join(csrc, "host", f) for f in os.listdir(os.path.join(csrc, "host")) if f.endswith(".h")]
    deps += [os.path.join(ROOT, "tests", "hostcheck", "cuda_shim.h"), os.path.join(ROOT, "include", "pbrt_r3_b200.h")]
    if not os.path.exists(HOSTCHECK_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOSTCHECK_LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DPT_HOSTCHECK", "-DPT_NCS=3", "-I", os.path.join(ROOT, "tests", "hostcheck"),
                               "-x", "c++", "-pthread", "-o", HOSTCHECK_LIB] + srcs + ["-ldl", "-lz"])
    return HOSTCHECK_LIB


@pytest.fixture(scope="module")
def HC(P):
    lib = build_hostcheck()
    os.environ["PBRT_R3_B200_DATA"] = os.path.join(ROOT, "pbrt-r3_b200", "data")
    saved = P.capi._lib
    P.capi._lib = P.capi.load(lib)
    yield P
    P.capi._lib = saved
    os.environ.pop("PBRT_R3_B200_DATA", None)
