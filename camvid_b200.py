"""Import shim: exposes the package directory `pytorch-camvid_b200/` (not a valid Python identifier) as the module
`camvid_b200`. `import camvid_b200` from the repository root (or with the root on sys.path) is the supported entry."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pytorch-camvid_b200")
_spec = importlib.util.spec_from_file_location("camvid_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["camvid_b200"] = _mod
_spec.loader.exec_module(_mod)
