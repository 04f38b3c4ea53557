"""Importable alias of the product package.

The product package directory is ``oldoceananigans.jl_b200/`` (the layout the build contract names); a dot
in a directory name cannot be imported, so this shim puts that directory on its ``__path__`` and re-exports
the public API.  All code lives in ``oldoceananigans.jl_b200/``.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "oldoceananigans.jl_b200")
__path__.append(_pkg_dir)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
