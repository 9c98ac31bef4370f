"""Import shim: the package directory is `ekf-slam-ml_b200/` (hyphenated, after the reference repo), which is
not a Python identifier.  `import ekf_slam_ml_b200` loads that directory as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ekf-slam-ml_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
