"""Registers the product package directory `open-headstage_b200/` (not a valid Python identifier) under the
importable name `open_headstage_b200`.  Used by tests/conftest.py, bench.py and __graft_entry__.py."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "open-headstage_b200")
PKG_NAME = "open_headstage_b200"


def load_package():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(PKG_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
