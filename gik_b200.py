"""Import shim: the package directory carries the repository's name, which is not a Python identifier
(`motion-planning-and-control-for-dual-manipulator-robot_b200/`); `import gik_b200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "motion-planning-and-control-for-dual-manipulator-robot_b200")
_spec = importlib.util.spec_from_file_location("gik_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gik_b200"] = _mod
_spec.loader.exec_module(_mod)
