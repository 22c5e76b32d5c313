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


def _source_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    for d in DEPS:
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """The library is current when the hash of its sources (recorded next to it at build time) matches — file times are
    not trusted: the snapshot that carries the prebuilt .so to the GPU box does not preserve them."""
    if not os.path.exists(LIB):
        return True
    if os.environ.get("OHS_LIB_OVERRIDE"):
        return False  # an explicitly chosen library (A/B experiments) is used as it is
    try:
        with open(LIB + ".srchash") as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        tmp = LIB + ".tmp.%d" % os.getpid()
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", tmp, *SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout + r.stderr)
        if r.returncode:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed building libohs_cuda.so")
        os.replace(tmp, LIB)  # atomic: concurrent ranks never see a half-written library
        with open(LIB + ".srchash", "w") as f:
            f.write(_source_hash())
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
