"""Import alias: the package lives in the directory `ftrl-ffm_b200/` (hyphenated, so it
cannot be named in an `import` statement).  `import ftrl_ffm_b200` loads that directory as
the package `ftrl_ffm_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ftrl-ffm_b200")
_spec = importlib.util.spec_from_file_location(
    "ftrl_ffm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ftrl_ffm_b200"] = _mod
_spec.loader.exec_module(_mod)
