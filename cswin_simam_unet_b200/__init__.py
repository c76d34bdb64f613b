"""Importable name of the ``cswin-simam-unet_b200/`` package directory.

The package directory carries the repository's (hyphenated) name, which Python cannot import directly.  This
stub loads that directory's ``__init__.py`` through the regular import machinery as the package
``cswin_simam_unet_b200`` (its ``__file__`` / ``__path__`` are the real ones, relative imports resolve there) and
puts it in ``sys.modules`` under this name.
"""
import importlib.util as _util
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cswin-simam-unet_b200")
_spec = _util.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"),
                                      submodule_search_locations=[_real])
_module = _util.module_from_spec(_spec)
_sys.modules[__name__] = _module
try:
    _spec.loader.exec_module(_module)
except BaseException:
    _sys.modules.pop(__name__, None)
    raise
