"""Import shim: the package directory is `arm-neon-opencv_b200/` (hyphens), so load it by path.

    import vacv_b200 as vacv
"""
import importlib.util
import os
import sys

_pkg = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arm-neon-opencv_b200")
_spec = importlib.util.spec_from_file_location("arm_neon_opencv_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["arm_neon_opencv_b200"] = _mod
_spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
