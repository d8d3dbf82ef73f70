"""Import shim: the package sources live in ``mast-multiphysics_b200/`` (not a valid Python
identifier), this makes them importable as ``mast_b200``."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "mast-multiphysics_b200"))

from .api import *  # noqa: E402,F401,F403
from . import mesh  # noqa: E402,F401
from . import partition  # noqa: E402,F401
