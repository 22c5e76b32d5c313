"""In-tree nvcc build of the C-ABI library (libohs_cuda.so) for sm_100a.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("OHS_LIB_OVERRIDE") or os.path.join(HERE, "libohs_cuda.so")  # override: A/B experiments only
CSRC = os.path.join(HERE, "csrc")
RENDER_SIZES = (128, 256, 512, 1024, 2048)   # transform sizes N = 2 * block; one object of csrc/ohs_render.cu per size
SOURCES = [os.path.join(CSRC, "ohs_api.cu"), os.path.join(CSRC, "ohs_render.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "ohs_kernels.cuh"), os.path.join(CSRC, "ohs_aux_kernels.cuh"), os.path.join(CSRC, "ohs_launch.h"),
                  os.path.join(ROOT, "include", "ohs.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # no -use_fast_math: denormals are kept (-ftz=false) and division/sqrt stay IEEE, as on the reference's CPU path
]
EXTRA_FLAGS = os.environ.get("OHS_NVCC_EXTRA", "").split()   # e.g. -DOHS_TRACE for the instrumented A/B library


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
    h.update(" ".join(NVCC_FLAGS + EXTRA_FLAGS).encode())
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


def _run(cmd, verbose):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if verbose or r.returncode:
        print(" ".join(cmd))
        print(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libohs_cuda.so")


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, extra_flags=()) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a: ohs_api.cu and one object of ohs_render.cu per transform size,
    compiled in parallel, linked into one shared library."""
    lib = out or LIB
    if not (force or out or is_stale()):
        return lib
    from concurrent.futures import ThreadPoolExecutor

    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build", "obj.%d" % os.getpid())
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + EXTRA_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else [])
    jobs = [([nvcc, *flags, "-c", SOURCES[0], "-o", os.path.join(objdir, "ohs_api.o")])]
    for n in RENDER_SIZES:
        jobs.append([nvcc, *flags, "-DOHS_RENDER_N=%d" % n, "-c", SOURCES[1], "-o", os.path.join(objdir, "ohs_render_%d.o" % n)])
    try:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
        tmp = lib + ".tmp.%d" % os.getpid()
        _run([nvcc, "-shared", "-o", tmp, *[j[-1] for j in jobs], "-ldl"], verbose)
        os.replace(tmp, lib)  # atomic: concurrent ranks never see a half-written library
        if not out:
            with open(LIB + ".srchash", "w") as f:
                f.write(_source_hash())
    finally:
        shutil.rmtree(objdir, ignore_errors=True)
    return lib


HOST_TEST_BIN = os.path.join(HERE, "host", "ref_unit_tests")
HOST_INPUTS_BIN = os.path.join(HERE, "host", "host_inputs_tool")


def build_host_tests(force: bool = False) -> str:
    """g++ build of the C++ mirror's replay of the reference's unit tests (links libohs_cuda.so by rpath)."""
    src = os.path.join(HERE, "host", "ref_unit_tests.cpp")
    deps = [src, os.path.join(HERE, "host", "dsp.hpp"), os.path.join(ROOT, "include", "ohs.h")]
    if force or not os.path.exists(HOST_TEST_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_TEST_BIN) for d in deps):
        build_library()
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_TEST_BIN, src, "-L" + HERE, "-lohs_cuda", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd, cwd=ROOT)
    return HOST_TEST_BIN


def build_host_inputs_tool(force: bool = False) -> str:
    """g++ build of host/host_inputs_tool.cpp: the C++ SOFA reader (zlib) and AutoEQ parser next to the hot path, plus a
    render mode that drives them into the engine through the C++ mirror objects."""
    src = os.path.join(HERE, "host", "host_inputs_tool.cpp")
    deps = [src] + [os.path.join(HERE, "host", f) for f in ("dsp.hpp", "sofa.hpp", "autoeq.hpp")] + [os.path.join(ROOT, "include", "ohs.h")]
    if force or not os.path.exists(HOST_INPUTS_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_INPUTS_BIN) for d in deps):
        build_library()
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_INPUTS_BIN, src, "-L" + HERE, "-lohs_cuda", "-lz", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd, cwd=ROOT)
    return HOST_INPUTS_BIN
