"""Import shim: makes the package directory ``neural-gauge-fields_b200/`` (not a valid Python identifier)
importable as ``ngf_b200``.  ``import ngf_b200`` executes this file, which loads the package's ``__init__`` under
the same module name, so ``ngf_b200.triplane`` etc. resolve as ordinary submodules."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "neural-gauge-fields_b200")
_spec = _ilu.spec_from_file_location("ngf_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["ngf_b200"] = _mod
_spec.loader.exec_module(_mod)
