#!/usr/bin/env python
"""Build the kernels of the working tree (or of a git revision) into open-headstage_b200/libohs_cuda_<tag>.so for A/B
runs with OHS_LIB_OVERRIDE (the library is git-ignored and travels to the GPU box with the snapshot).

usage: ab_build.py <tag> [--rev REV] [extra nvcc flags, e.g. -DOHS_TRACE -DOHS_EQ_WEIGHT=3]
"""
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    tag = args.pop(0)
    rev = None
    if args[:1] == ["--rev"]:
        rev = args[1]
        args = args[2:]
    out = os.path.join(ROOT, "open-headstage_b200", "libohs_cuda_%s.so" % tag)
    tree = ROOT
    tmp = None
    if rev:
        tmp = tempfile.TemporaryDirectory()
        tree = tmp.name
        subprocess.check_call("git -C %s archive %s open-headstage_b200 include | tar -x -C %s" % (ROOT, rev, tree), shell=True)
    spec = importlib.util.spec_from_file_location("_ab_build", os.path.join(tree, "open-headstage_b200", "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build_library(force=True, out=out, extra_flags=args)
    except TypeError:  # revisions before the per-size translation units: one nvcc call over ohs_api.cu
        subprocess.check_call([mod.nvcc_path(), *mod.NVCC_FLAGS, *args, "-o", out, *mod.SOURCES], cwd=tree)
    print("built", out, "from", rev or "the working tree", " ".join(args))


if __name__ == "__main__":
    main()
