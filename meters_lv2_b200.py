"""Import shim: `import meters_lv2_b200` loads the package that lives in the directory
`meters.lv2_b200/` (the dot makes that directory name unimportable by the normal machinery)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meters.lv2_b200")
_spec = importlib.util.spec_from_file_location(
    "meters_lv2_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["meters_lv2_b200"] = _mod
_spec.loader.exec_module(_mod)
