"""In-tree nvcc build of the C-ABI library (libohs_cuda.so) for sm_100a.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("OHS_LIB_OVERRIDE") or os.path.join(HERE, "libohs_cuda.so")  # override: A/B experiments only
SOURCES = [os.path.join(HERE, "csrc", "ohs_api.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "ohs_kernels.cuh"), os.path.join(ROOT, "include", "ohs.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # no -use_fast_math: denormals are kept (-ftz=false) and division/sqrt stay IEEE, as on the reference's CPU path
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libohs_cuda.so cannot be built (there is no CPU fallback)")
    return p


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building libohs_cuda.so")
    return LIB


HOST_TEST_BIN = os.path.join(HERE, "host", "ref_unit_tests")


def build_host_tests(force: bool = False) -> str:
    """g++ build of the C++ mirror's replay of the reference's unit tests (links libohs_cuda.so by rpath)."""
    src = os.path.join(HERE, "host", "ref_unit_tests.cpp")
    deps = [src, os.path.join(HERE, "host", "dsp.hpp"), os.path.join(ROOT, "include", "ohs.h")]
    if force or not os.path.exists(HOST_TEST_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_TEST_BIN) for d in deps):
        build_library()
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_TEST_BIN, src, "-L" + HERE, "-lohs_cuda", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd, cwd=ROOT)
    return HOST_TEST_BIN
