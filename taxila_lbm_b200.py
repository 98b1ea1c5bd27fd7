"""Import shim: the package directory is `taxila-lbm_b200/` (a hyphen is not a
valid module name), so `import taxila_lbm_b200` loads that directory as a package
under this name."""
import importlib.util
import sys
from pathlib import Path

_root = Path(__file__).resolve().parent / "taxila-lbm_b200"
_spec = importlib.util.spec_from_file_location(
    "taxila_lbm_b200", _root / "__init__.py", submodule_search_locations=[str(_root)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["taxila_lbm_b200"] = _mod
_spec.loader.exec_module(_mod)
