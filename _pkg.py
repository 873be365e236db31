"""Loader for the host-side package, whose directory name (``q-gcm_b200``) is not a
valid Python identifier.  ``load()`` registers it as module ``qgcm_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "q-gcm_b200")


def load():
    if "qgcm_b200" in sys.modules:
        return sys.modules["qgcm_b200"]
    spec = importlib.util.spec_from_file_location(
        "qgcm_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["qgcm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
