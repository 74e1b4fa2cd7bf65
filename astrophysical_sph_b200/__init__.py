"""Importable alias of the package directory `astrophysical-sph_b200/` (a hyphen cannot be imported).

Replaces itself in sys.modules with the real package, whose submodules live in the hyphenated directory.
"""
import os
import sys
from importlib.util import module_from_spec, spec_from_file_location

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "astrophysical-sph_b200")
_spec = spec_from_file_location(__name__, os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
