"""Makes the hyphenated product directory `del-fem_b200/` importable as the package `del_fem_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "del_fem_b200" in sys.modules:
        return sys.modules["del_fem_b200"]
    pkg_dir = os.path.join(ROOT, "del-fem_b200")
    spec = importlib.util.spec_from_file_location("del_fem_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["del_fem_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
