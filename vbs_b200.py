"""Import shim: the package directory is named ``vision-basedsensor_b200`` (not a
valid Python identifier), so ``import vbs_b200`` loads it under this alias.

After this module runs, ``sys.modules['vbs_b200']`` IS the package (with its
``__path__``), so ``import vbs_b200.pipeline`` etc. resolve normally.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vision-basedsensor_b200")
_spec = _ilu.spec_from_file_location(
    "vbs_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["vbs_b200"] = _mod
_spec.loader.exec_module(_mod)
