"""Import shim: makes the package in ``gp-mpc_b200/`` importable as ``gp_mpc_b200``
(a hyphen is not a valid module name)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'gp-mpc_b200')
_spec = _u.spec_from_file_location('gp_mpc_b200', _os.path.join(_dir, '__init__.py'),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules['gp_mpc_b200'] = _mod
_spec.loader.exec_module(_mod)
