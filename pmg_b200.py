"""Import shim: the package directory is named `parallel-geometric-multigrid-for-poisson-problem_b200`
(hyphens are not importable), so `import pmg_b200` loads it by path."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "parallel-geometric-multigrid-for-poisson-problem_b200")
_spec = importlib.util.spec_from_file_location("_pmg_b200_pkg", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["_pmg_b200_pkg"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
